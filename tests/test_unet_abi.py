"""not-gpu tests of the U-Net entry points of the model-level C ABI (csrc/tpz_unet.cu: tpz_unet_create / tpz_unet2d_forward /
tpz_unet3d_forward).  The plan builder and the launch sequence are C++; the kernels need a GPU.  The library's launch hook hands
every launch the C++ code would make to tests/sim_backend.py (the CPU simulation of the kernel semantics, interpreting the very
argument blocks the device entry points receive), on a handle whose packed buffers stay in host memory.  Checked here:
  * packed fp16 weights, biases, k-block tables and static argument blocks equal the Python engine's plans bit for bit;
  * the network output equals the Python engine's (same simulation) bit for bit, and the reference goldens at the usual tolerance;
  * the workspace bound holds for every buffer a launch touches; out-of-range weights and bad geometry are refused."""
import ctypes as C
import traceback

import numpy as np
import pytest
import torch

from common import gold, weights_of, seeded_state, rel_err
from common_shapes import unet_shapes
import sim_backend

from topaz_b200 import _lib, engine, ops
from topaz_b200.model_abi import UnetModel, WeightRangeError

_CT = {torch.float16: (C.c_uint16, np.float16), torch.float32: (C.c_float, np.float32)}


def _view(ptr, shape, dtype):
    """torch tensor aliasing host memory at `ptr`"""
    n = int(np.prod(shape))
    ct, npd = _CT[dtype]
    arr = np.ctypeslib.as_array((ct * n).from_address(ptr))
    return torch.from_numpy(arr.view(npd).reshape(shape))


class SimHook:
    """tpz_launch_hook: runs each launch of the C++ U-Net on the CPU simulation; records (op, touched byte ranges)"""

    def __init__(self):
        self.ops = []
        self.touched = []
        self.error = None
        self.fn = _lib.LAUNCH_HOOK(self._call)

    def _t(self, ptr, shape, dtype):
        self.touched.append((ptr, int(np.prod(shape)) * (2 if dtype == torch.float16 else 4)))
        return _view(ptr, shape, dtype)

    def _call(self, user, op, args):
        try:
            self.ops.append(op)
            self._run(op, args)
            return 0
        except Exception:                                   # an exception must not unwind through the C frames
            self.error = traceback.format_exc()
            return 99

    def _run(self, op, args):
        f16, f32 = torch.float16, torch.float32
        if op == 5:
            return self._tc_conv(_lib.TpzTcConvArgs.from_address(args))
        o = _lib.TpzOpArgs.from_address(args)
        p, i, f = o.p, o.i, o.f
        rng = lambda k: _view(p[k], (2,), f32) if p[k] else None
        if op == 0:
            x = _view(p[0], (o.n,), f32)
            self._t(p[1], (2,), f32).copy_(sim_backend.range_scale(x))
        elif op == 1:
            B, H, W, Cp, k, pad, pool = i[0:7]
            w = _view(p[1], ((k * k + 63) // 64, Cp, 64), f16)
            y = sim_backend.conv_first_tc(_view(p[0], (B, H, W), f32), w, _view(p[2], (Cp,), f32), k, pad, f[0], pool=bool(pool), rng=rng(4))
            self._t(p[3], tuple(y.shape), f16).copy_(y)
        elif op == 2:
            N, H, W, k, pad, ld, out_lo = i[0:7]
            assert out_lo in (0, ld)
            y = sim_backend.im2col_first(_view(p[0], (N, H, W), f32), k, pad, ld, rng=rng(2), split=out_lo > 0)
            self._t(p[1], tuple(y.shape), f16).copy_(y)
        elif op == 3:
            N, D, H, W, k, pad, ld, out_lo = i[0:8]
            assert out_lo in (0, ld) and pad == k // 2
            y = sim_backend.im2col3d_first(_view(p[0], (N, D, H, W), f32), k, ld, rng=rng(2), split=out_lo > 0)
            self._t(p[1], tuple(y.shape), f16).copy_(y)
        elif op == 4:
            N, D, H, W, Co, kd, kh, kw, dil, pad, pool, out_ld, out_lo = i[0:13]
            assert pool == 1 and (out_lo == 0 or 2 * out_lo == out_ld)
            y = sim_backend.conv_first(_view(p[0], (N, D, H, W), f32), _view(p[1], (Co, kd, kh, kw), f32), _view(p[2], (Co,), f32), dil, pad,
                                       f[0], out_lo if out_lo else out_ld, rng=rng(4), split=out_lo > 0)
            self._t(p[3], tuple(y.shape), f16).copy_(y)
        elif op == 6:
            N, D, H, W, Cc, ld, dims, out_ld, lo_off = i[0:9]
            assert ld == out_ld and ((lo_off == 0 and Cc == ld) or (lo_off == Cc and 2 * Cc == ld))
            y = sim_backend.maxpool2(_view(p[0], (N, D, H, W, ld), f16), dims, split=lo_off > 0)
            self._t(p[1], tuple(y.shape), f16).copy_(y)
        elif op == 7:
            N, D, H, W, Cc, ld, Do, Ho, Wo, out_ld, coff = i[0:11]
            assert Cc == ld == out_ld and coff == 0
            y = sim_backend.upsample_nearest(_view(p[0], (N, D, H, W, ld), f16), (Do, Ho, Wo))
            self._t(p[1], tuple(y.shape), f16).copy_(y)
        elif op == 8:
            N, D, H, W, Cc, ld, kd, kh, kw, dil, pad = i[0:11]
            stats = _view(p[2], (2,), f32) if p[2] else None
            y = sim_backend.conv_last(_view(p[0], (N, D, H, W, ld), f16), Cc, _view(p[1], (kd * kh * kw, Cc), f32), f[0], (kd, kh, kw), dil,
                                      pad, stats=stats, out_scale=f[1], out_shift=f[2], rng=rng(4))
            _view(p[3], (N, D, H, W), f32).copy_(y)
        else:
            raise AssertionError(f'unknown op {op}')

    def _tc_conv(self, a):
        f16, f32 = torch.float16, torch.float32
        plan = plan_of_args(a)
        srcs = [self._t(a.src[s].ptr, (a.src[s].N, a.src[s].D, a.src[s].H, a.src[s].W, a.src[s].ld), f16) for s in range(a.nsrc)]
        shape = (a.N, a.Do, a.Ho, a.Wo)
        assert not a.res and not a.oscale and a.out_coff == 0 and a.out_lo in (0, a.Co)
        assert a.out_ld == (2 * a.Co if a.out_lo else a.Co) or not a.out
        plan.split_out = a.out_lo > 0
        out = self._t(a.out, shape + (a.out_ld,), f16) if a.out else None
        dot_out = _view(a.dot_out, shape, f32) if a.dot_out else None
        dot_affine = _view(a.dot_affine, (2,), f32) if a.dot_affine else None
        rng = _view(a.range, (2,), f32) if a.range else None
        sim_backend.tc_conv(plan, srcs, shape, out=out, dot_out=dot_out, dot_affine=dot_affine, rng=rng)


def plan_of_args(a) -> ops.TcConvPlan:
    """the ops.TcConvPlan an argument block describes (weights / bias alias the block's host buffers)"""
    f16, f32 = torch.float16, torch.float32
    n = a.nsrc
    return ops.TcConvPlan(
        KC=a.KC, Co=a.Co, kblocks=[(a.kb[j].dx, a.kb[j].dy, a.kb[j].dz, a.kb[j].c0, a.kb[j].src) for j in range(a.nkb)],
        orgs=[tuple(a.src[s].org) for s in range(n)], c_stores=[a.src[s].C for s in range(n)],
        tapgrids=[(a.src[s].kw, a.src[s].kh) for s in range(n)], lattice=a.lattice,
        weights=_view(a.weights, (a.nkb, a.Co, a.KC), f16), bias=_view(a.bias, (a.Co,), f32), neg_slope=a.neg_slope,
        lats=[a.src[s].lat for s in range(n)], phases=[not a.src[s].no_phase for s in range(n)], phase_sel=a.phase_sel,
        lat_zs=[a.src[s].lat_z for s in range(n)], lattice_z=a.lattice_z, phase_z=a.phase_z,
        dot_w=_view(a.dot_w, (a.Co,), f32) if a.dot_w else None, dot_b=a.dot_b, TW=a.TW, TH=a.TH)


def run_c(model, x, stats=None, check_bounds=True):
    """forward of the C++ U-Net on host tensors under the simulation hook -> (y, ops launched)"""
    um = UnetModel(model, host=True, precision=engine.PRECISION)
    hook = SimHook()
    lib = _lib.lib()
    shape = (x.shape[0], 1, x.shape[2], x.shape[3]) if um.dims == 2 else (x.shape[0],) + tuple(x.shape[2:])
    need = um.workspace_bytes(shape)
    raw = torch.zeros(need + 512, dtype=torch.uint8)
    off = (-raw.data_ptr()) % 256
    ws = raw[off:off + need]
    raw[off + need:] = 0xA5                                   # canary behind the workspace
    lib.tpz_unet_set_launch_hook(hook.fn, None)
    try:
        y = um.forward(x, denorm_stats=stats, workspace=ws)
    finally:
        lib.tpz_unet_set_launch_hook(C.cast(None, _lib.LAUNCH_HOOK), None)
    assert hook.error is None, hook.error
    assert um.launch_count(shape) == len(hook.ops) + 1        # the range scale is two kernels
    if check_bounds:
        lo, hi = ws.data_ptr() + 1024, ws.data_ptr() + need
        for ptr, nbytes in hook.touched:
            if ptr == ws.data_ptr():                           # the range pair in the workspace header
                continue
            assert lo <= ptr and ptr + nbytes <= hi and ptr % 256 == 0, (ptr - ws.data_ptr(), nbytes, need)
        assert bool((raw[off + need:] == 0xA5).all())
    um.close()
    return y, hook.ops


def run_py(model, x, stats=None):
    with sim_backend.patched(), torch.no_grad():
        return engine.unet_forward(model, x, stats)


def _load(model, sd):
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    model.eval()
    return model


@pytest.fixture
def fast_precision():
    saved = engine.PRECISION
    engine.PRECISION = 'fast'
    yield
    engine.PRECISION = saved


def _static_fields(a):
    srcs = [(a.src[s].C, tuple(a.src[s].org), a.src[s].kw, a.src[s].kh, a.src[s].lat, a.src[s].no_phase, a.src[s].lat_z) for s in range(a.nsrc)]
    kbs = [(a.kb[j].dx, a.kb[j].dy, a.kb[j].dz, a.kb[j].c0, a.kb[j].src) for j in range(a.nkb)]
    return (a.nsrc, srcs, a.KC, a.nkb, kbs, a.Co, a.TW, a.TH, a.lattice, a.phase_sel, a.lattice_z, a.phase_z, a.neg_slope, a.dot_b)


def _assert_same_plan(c_args, n_elems, py_plan, what):
    pa = ops._static_tc_args(py_plan)
    pa.dot_b = py_plan.dot_b
    assert _static_fields(c_args) == _static_fields(pa), what
    assert n_elems == py_plan.weights.numel(), what
    cw = _view(c_args.weights, tuple(py_plan.weights.shape), torch.float16)
    assert torch.equal(cw.view(torch.int16), py_plan.weights.view(torch.int16)), what          # bit-identical fp16 blocks
    assert torch.equal(_view(c_args.bias, (py_plan.Co,), torch.float32), py_plan.bias), what
    assert py_plan.oscale is None


def _compare_all_plans(model):
    py = engine._build_unet_plan(model, 'cpu')
    um = UnetModel(model, host=True, precision=engine.PRECISION)
    n = 0
    if py['first_tc'] is not None and py['first_fused'] is None:
        _assert_same_plan(*um.plan(0), py['first_tc']['plan'], 'first'); n += 1
    for i, e in enumerate(py['enc']):
        _assert_same_plan(*um.plan(1, i), e['plan'], f'enc{i + 2}'); n += 1
    for l, d in py['dec'].items():
        _assert_same_plan(*um.plan(2, l), d['a'], f'dec{l}.0'); n += 1
        _assert_same_plan(*um.plan(3, l), d['b'], f'dec{l}.2'); n += 1
        for ph, pl in enumerate(d['up2']):
            _assert_same_plan(*um.plan(4, l, ph), pl, f'dec{l}.0 phase {ph}'); n += 1
    _assert_same_plan(*um.plan(5), py['dec'][1]['last_tc'], 'dec1.4')
    a, _ = um.plan(5)
    assert float(_view(a.weights, (1,), torch.float16)[0]) == float(py['dec'][1]['last_tc'].weights.reshape(-1)[0])
    um.close()
    return n


def test_unet2d_pretrained_plans_and_forward_match_python_engine():
    from topaz_b200.denoising.models import UDenoiseNet
    g = gold('unet_pretrained')
    m = _load(UDenoiseNet(base_width=11, top_width=5), weights_of(g))
    assert _compare_all_plans(m) == 5 + 5 * 2 + 5 * 4
    for xk, yk in (('x', 'y'), ('xo', 'yo')):                 # 96x128: every level an exact 2x (fused up-sampling); 95x77: none
        x = torch.from_numpy(g[xk])
        yc, launched = run_c(m, x)
        assert torch.equal(yc, run_py(m, x)), xk
        mx, l2 = rel_err(yc.numpy(), g[yk])
        assert mx < 2e-3 and l2 < 2e-3, (mx, l2)
        assert (7 in launched) == (xk == 'xo')                # materialised up-sampling only where the sizes are not exact doubles
        assert launched[0] == 0 and launched[1] == 1 and launched[-1] == 8


def test_unet2d_small_pretrained_and_denormalise_epilogue():
    from topaz_b200.denoising.models import UDenoiseNetSmall
    g = gold('unet_small_pretrained')
    m = _load(UDenoiseNetSmall(width=11, top_width=5), weights_of(g))
    assert _compare_all_plans(m) == 3 + 3 * 2 + 3 * 4
    x = torch.from_numpy(g['xo'])
    yc, _ = run_c(m, x)
    assert torch.equal(yc, run_py(m, x))
    mx, l2 = rel_err(yc.numpy(), g['yo'])
    assert mx < 2e-3 and l2 < 2e-3, (mx, l2)
    stats = torch.tensor([10.0, 3.0])
    yd, _ = run_c(m, x, stats=stats)
    assert torch.equal(yd, run_py(m, x, stats))
    assert torch.allclose(yd, yc * 3.0 + 10.0, rtol=1e-6, atol=1e-6)


def test_unet2d_seeded_batch_and_large_magnitude_input():
    from topaz_b200.denoising.models import UDenoiseNet
    g = gold('unet_seeded_nf16')
    m = _load(UDenoiseNet(nf=16, base_width=7, top_width=3), seeded_state(unet_shapes(16, 7, 3, 2), int(g['seed'])))
    _compare_all_plans(m)
    x = torch.from_numpy(g['x'])                              # batch of 2
    yc, _ = run_c(m, x)
    assert torch.equal(yc, run_py(m, x))
    mx, l2 = rel_err(yc.numpy(), g['y'])
    assert mx < 3e-3 and l2 < 3e-3, (mx, l2)
    xl = x * 3.0e5                                            # un-normalised input: the range scale keeps fp16 in range
    yl, _ = run_c(m, xl)
    assert torch.isfinite(yl).all() and torch.equal(yl, run_py(m, xl))


@pytest.mark.parametrize('base,top,mode_op', [(9, 3, 2), (13, 7, 4)])
def test_unet2d_other_first_layers(base, top, mode_op):
    """9x9: no fused first-layer kernel -> im2col + GEMM; 13x13: more than 128 taps -> fp32 CUDA-core first layer.  7x7 top: the
    Cout = 1 tail on the tensor-core path (no tiled CUDA-core kernel for 7x7)."""
    from topaz_b200.denoising.models import UDenoiseNetSmall
    torch.manual_seed(base)
    m = UDenoiseNetSmall(nf=16, width=base, top_width=top).eval()
    _compare_all_plans(m)
    x = torch.randn(1, 1, 40, 56)
    yc, launched = run_c(m, x)
    assert torch.equal(yc, run_py(m, x))
    assert launched[1] == mode_op
    assert (launched[-1] == 5) == (top == 7)


def test_unet3d_seeded_fast_precision(fast_precision):
    from topaz_b200.denoising.models import UDenoiseNet3D
    g = gold('unet3d_seeded')
    m = _load(UDenoiseNet3D(nf=48, base_width=7, top_width=3), seeded_state(unet_shapes(48, 7, 3, 3), int(g['seed'])))
    assert _compare_all_plans(m) == 1 + 5 + 5 * 2 + 5 * 8
    x = torch.from_numpy(g['x'])
    yc, launched = run_c(m, x)
    assert torch.equal(yc, run_py(m, x))
    mx, l2 = rel_err(yc.numpy(), g['y'])
    assert mx < 1e-2 and l2 < 1e-2, (mx, l2)                  # fast precision on the 3-D net (auto mode splits its last four convs)
    assert launched.count(3) == 1 and launched[-1] == 8
    x2 = torch.randn(1, 1, 32, 40, 36)                       # 36 / 2 / 2 = 9: odd level -> materialised up-sampling in 3-D
    y2, launched = run_c(m, x2)
    assert 7 in launched and torch.equal(y2, run_py(m, x2))


def test_unet_c_model_refuses_what_it_cannot_run():
    from topaz_b200.denoising.models import UDenoiseNetSmall
    torch.manual_seed(0)
    m = UDenoiseNetSmall(nf=16, width=7, top_width=3).eval()
    um = UnetModel(m, host=True)
    assert um.workspace_bytes((1, 1, 64, 64)) > 0
    with pytest.raises(RuntimeError):
        um.workspace_bytes((1, 1, 4, 64))                     # smaller than three pooling stages allow
    with pytest.raises(RuntimeError):                         # a host handle never launches on the device
        um.forward(torch.zeros(1, 1, 16, 16))
    um.close()
    with torch.no_grad():
        m.dec2[0].weight[3] *= 1e7                            # a row beyond the fp16 range needs the row-scaled (Python) plans
    with pytest.raises(WeightRangeError):
        UnetModel(m, host=True)
    with torch.no_grad():
        m.dec2[0].weight[3, 0, 0, 0] = float('nan')
    with pytest.raises(RuntimeError, match='non-finite'):
        UnetModel(m, host=True)


def test_engine_routes_unet_forward_through_the_c_handle(monkeypatch, fast_precision):
    """engine.unet_forward with TPZ_UNET_ENGINE=c: the handle is cached per parameter state and precision, rebuilt when a weight
    changes, and left aside (Python plans) under a non-default kernel selection."""
    from topaz_b200 import model_abi
    from topaz_b200.denoising.models import UDenoiseNetSmall, UDenoiseNet3D
    torch.manual_seed(3)
    m = UDenoiseNetSmall(nf=16, width=7, top_width=3).eval()
    x = torch.randn(1, 1, 32, 48)
    y_py = run_py(m, x)
    class HostUnetModel(UnetModel):
        def __init__(self, model, precision='fast'):
            super().__init__(model, host=True, precision=precision)
    monkeypatch.setattr(model_abi, 'UnetModel', HostUnetModel)
    monkeypatch.setattr(engine, 'UNET_ENGINE', 'c')
    hook = SimHook()
    lib = _lib.lib()
    lib.tpz_unet_set_launch_hook(hook.fn, None)
    try:
        with sim_backend.patched(), torch.no_grad():
            y_c = engine.unet_forward(m, x)
            h0 = m.__dict__['_tpz_plans']['unet_c'][1]
            assert h0 is not None and torch.equal(y_c, y_py) and y_c.shape == x.shape
            engine.unet_forward(m, x)
            assert m.__dict__['_tpz_plans']['unet_c'][1] is h0                 # cached
            m.enc2[0].weight.mul_(0.5)                                          # in-place update -> new handle, new result
            y2 = engine.unet_forward(m, x)
            assert m.__dict__['_tpz_plans']['unet_c'][1] is not h0
            assert torch.equal(y2, run_py(m, x)) and not torch.equal(y2, y_c)
            with pytest.raises(ValueError):
                engine.unet_forward(m, x[0])
            engine.PRECISION = 'strict'                                         # a new precision is a new handle
            ys = engine.unet_forward(m, x)
            assert m.__dict__['_tpz_plans']['unet_c'][1] is not None
            assert torch.equal(ys, run_py(m, x)) and not torch.equal(ys, y2)
            engine.PRECISION = 'auto'
            m3 = UDenoiseNet3D(nf=16, base_width=3, top_width=3).eval()
            assert engine._unet_c_model(m3, 'k') is not None
            monkeypatch.setattr(engine, 'UP2_FUSED', False)                     # non-default kernel selection: Python plans
            n_ops = len(hook.ops)
            yu = engine.unet_forward(m, x)
            assert m.__dict__['_tpz_plans']['unet_c'][1] is None and len(hook.ops) == n_ops and torch.isfinite(yu).all()
    finally:
        lib.tpz_unet_set_launch_hook(C.cast(None, _lib.LAUNCH_HOOK), None)
    assert hook.error is None, hook.error


_DYNAMIC = ('N', 'Do', 'Ho', 'Wo', 'out_ld', 'out_coff', 'out_lo', 'dot_b', 'res_ld', 'res_D', 'res_H', 'res_W')
_POINTERS = ('res', 'res_scale', 'out', 'dot_w', 'dot_out', 'dot_affine', 'oscale', 'range')


def _launch_fields(a):
    """everything a tpz_tc_conv launch passes except addresses: static block, per-launch geometry, which pointers are set"""
    srcs = [(a.src[s].N, a.src[s].D, a.src[s].H, a.src[s].W, a.src[s].ld, bool(a.src[s].ptr)) for s in range(a.nsrc)]
    return (_static_fields(a), srcs, tuple(getattr(a, k) for k in _DYNAMIC), tuple(a.res_org), tuple(bool(getattr(a, k)) for k in _POINTERS))


@pytest.mark.parametrize('dims', [2, 3])
def test_every_tc_conv_argument_block_equals_the_python_engines(dims, fast_precision, monkeypatch):
    """The argument block of EVERY tensor-core launch -- the struct tpz_tc_conv receives -- field by field against the one
    ops.fill_tc_args builds for the Python engine's launch at the same position (addresses excepted)."""
    from topaz_b200.denoising.models import UDenoiseNetSmall, UDenoiseNet3D
    torch.manual_seed(11)
    if dims == 2:
        m, x = UDenoiseNetSmall(nf=16, width=7, top_width=3).eval(), torch.randn(2, 1, 40, 52)     # 52/4 = 13: one odd level
    else:
        m, x = UDenoiseNet3D(nf=16, base_width=3, top_width=3).eval(), torch.randn(1, 1, 32, 32, 64)
    py_blocks = []
    with sim_backend.patched(), torch.no_grad():
        sim_tc = ops.tc_conv

        def recording_tc_conv(plan, srcs, out_shape, out=None, res=None, res_org=(0, 0, 0), dot_out=None, out_coff=0, tag=None,
                              dot_affine=None, rng=None):
            a = ops.fill_tc_args(plan, srcs, out_shape, out, res, res_org, dot_out, out_coff, dot_affine, rng)
            py_blocks.append(_launch_fields(a))
            return sim_tc(plan, srcs, out_shape, out=out, res=res, res_org=res_org, dot_out=dot_out, out_coff=out_coff, tag=tag,
                          dot_affine=dot_affine, rng=rng)
        monkeypatch.setattr(ops, 'tc_conv', recording_tc_conv)
        y_py = engine.unet_forward(m, x)
    c_blocks = []
    real_tc = SimHook._tc_conv

    def recording_hook_tc(self, a):
        c_blocks.append(_launch_fields(a))
        return real_tc(self, a)
    monkeypatch.setattr(SimHook, '_tc_conv', recording_hook_tc)
    y_c, _ = run_c(m, x)
    assert torch.equal(y_c, y_py)
    assert len(c_blocks) == len(py_blocks) and len(c_blocks) > 10
    for j, (cb, pb) in enumerate(zip(c_blocks, py_blocks)):
        assert cb == pb, f'launch {j}'


@pytest.fixture
def precision():
    saved = engine.PRECISION

    def set_(p):
        engine.PRECISION = p
    yield set_
    engine.PRECISION = saved


@pytest.mark.parametrize('mode', ['strict', 'auto'])
def test_unet3d_split_operand_layers(mode, precision):
    """strict: every layer with (hi, lo) operands; auto (the 3-D default): the last four convolutions.  Plans (tripled k-blocks,
    doubled source channels), the (hi, lo) stores of convs / first layer / im2col / max-pool, and the Cout = 1 tail on split inputs."""
    from topaz_b200.denoising.models import UDenoiseNet3D
    precision(mode)
    g = gold('unet3d_seeded')
    m = _load(UDenoiseNet3D(nf=48, base_width=7, top_width=3), seeded_state(unet_shapes(48, 7, 3, 3), int(g['seed'])))
    assert _compare_all_plans(m) == (0 if mode == 'strict' else 1) + 5 + 5 * 2 + 5 * 8
    x = torch.from_numpy(g['x'])
    yc, launched = run_c(m, x)
    assert torch.equal(yc, run_py(m, x))
    mx, l2 = rel_err(yc.numpy(), g['y'])
    assert mx < 2e-3 and l2 < 2e-3, (mx, l2)
    assert launched[1] == (4 if mode == 'strict' else 2) and launched[-1] == 8       # strict: fp32 first layer; tail on the CUDA cores
    x2 = torch.randn(1, 1, 32, 40, 36, generator=torch.Generator().manual_seed(2))   # odd level: materialised (hi, lo) up-sampling
    y2, _ = run_c(m, x2)
    assert torch.equal(y2, run_py(m, x2))


@pytest.mark.parametrize('top', [3, 7])
def test_unet2d_strict_precision(top, precision):
    """2-D strict: fp32 first layer with a (hi, lo) output, split pooling, split raw-image taps; 3x3 top -> the tail runs on the
    tensor-core plan with split operands (the tiled CUDA-core tail takes plain fp16 only); 7x7 top likewise."""
    from topaz_b200.denoising.models import UDenoiseNetSmall
    precision('strict')
    torch.manual_seed(top)
    m = UDenoiseNetSmall(nf=16, width=7, top_width=top).eval()
    _compare_all_plans(m)
    x = torch.randn(2, 1, 40, 52)
    yc, launched = run_c(m, x)
    assert torch.equal(yc, run_py(m, x))
    assert launched[1] == 4 and launched[-1] == 5
    precision('fast')
    yf, _ = run_c(m, x)
    with torch.no_grad():
        ref = torch_reference(m, x)
    e_strict, e_fast = float((yc - ref).abs().max()), float((yf - ref).abs().max())
    assert e_strict < 0.2 * e_fast, (e_strict, e_fast)        # the split operands buy > 5x accuracy on the same network


def torch_reference(m, x):
    """UDenoiseNetSmall.forward of the reference (denoising/models.py:200-244) in plain torch fp32 on the drop-in's parameters"""
    import torch.nn.functional as F
    act = lambda t: F.leaky_relu(t, 0.1)
    conv = lambda c, t: F.conv2d(t, c.weight, c.bias, padding=c.weight.shape[-1] // 2)
    p1 = F.max_pool2d(act(conv(m.enc1[0], x)), 2)
    p2 = F.max_pool2d(act(conv(m.enc2[0], p1)), 2)
    p3 = F.max_pool2d(act(conv(m.enc3[0], p2)), 2)
    h = act(conv(m.enc4[0], p3))
    for dec, skip in ((m.dec3, p2), (m.dec2, p1)):
        h = torch.cat([F.interpolate(h, size=skip.shape[2:], mode='nearest'), skip], 1)
        h = act(conv(dec[2], act(conv(dec[0], h))))
    h = torch.cat([F.interpolate(h, size=x.shape[2:], mode='nearest'), x], 1)
    return conv(m.dec1[4], act(conv(m.dec1[2], act(conv(m.dec1[0], h)))))


@pytest.mark.parametrize('seed', range(10))
def test_random_architectures_and_patch_sizes(seed, precision):
    """Randomised depth / width / kernel sizes / patch shapes / precision: the C++ handle and the Python engine must agree bit for
    bit on every one (also at the smallest patch the pooling stages allow, where whole levels are a single pixel)."""
    from topaz_b200.denoising.models import _UNetBase
    rs = np.random.RandomState(1000 + seed)
    dims = 3 if seed % 3 == 2 else 2

    class Net(_UNetBase):
        _dims = dims

        def __init__(self, nf, base, top, depth):
            super().__init__()
            self._build(nf, base, top, depth)
    depth = int(rs.choice([3, 4] if dims == 3 else [3, 4, 5, 6]))     # (_build's dec1 assumes a 2*nf-channel decoder level below it)
    nf = int(rs.choice([8, 16, 24] if dims == 3 else [8, 16, 40, 48]))
    base = int(rs.choice([3, 5] if dims == 3 else [3, 5, 7, 9, 11]))
    top = int(rs.choice([3] if dims == 3 else [3, 5]))
    precision(['fast', 'auto', 'strict'][seed % 3] if dims == 3 else ['fast', 'strict'][seed % 2])
    torch.manual_seed(seed)
    m = Net(nf, base, top, depth).eval()
    div = 1 << (depth - 1)
    shapes = [tuple(int(div * rs.randint(1, 4) + rs.randint(0, div)) for _ in range(dims)), (div,) * dims]
    for sp in shapes:
        x = torch.randn((1 + seed % 2, 1) + sp, generator=torch.Generator().manual_seed(seed))
        yc, _ = run_c(m, x)
        assert torch.equal(yc, run_py(m, x)), (seed, dims, depth, nf, base, top, sp, engine.PRECISION)
    _compare_all_plans(m)


import contextlib
import os

from test_dropin_reference_cli import aliased      # noqa: F401  (fixture: the reference package aliased onto the drop-in modules)


@contextlib.contextmanager
def c_engine_under_hook(monkeypatch):
    """engine.unet_forward -> host-memory C handle -> launch hook -> CPU simulation, for code that only knows the nn.Modules"""
    from topaz_b200 import model_abi

    class HostUnetModel(UnetModel):
        def __init__(self, model, precision='fast'):
            super().__init__(model, host=True, precision=precision)
    monkeypatch.setattr(model_abi, 'UnetModel', HostUnetModel)
    monkeypatch.setattr(engine, 'UNET_ENGINE', 'c')
    hook = SimHook()
    _lib.lib().tpz_unet_set_launch_hook(hook.fn, None)
    try:
        yield hook
    finally:
        _lib.lib().tpz_unet_set_launch_hook(C.cast(None, _lib.LAUNCH_HOOK), None)
    assert hook.error is None, hook.error


def test_existing_unet_suite_through_the_c_handle(monkeypatch):
    """tests/test_host_logic.py's U-Net goldens (pretrained 2-D, seeded 2-D, seeded 3-D in the default `auto` precision) with every
    forward going through tpz_unet2d_forward / tpz_unet3d_forward"""
    import test_host_logic as H
    with c_engine_under_hook(monkeypatch) as hook:
        H.test_unet_sim_pretrained_and_seeded()
    assert hook.ops.count(0) == 4                 # four forwards, each opened by its range scale


@pytest.mark.skipif(not os.path.isdir('/root/reference/topaz'), reason='reference checkout not present')
def test_reference_pipelines_through_the_c_handle(aliased, tmp_path, monkeypatch):   # noqa: F811
    """The UNMODIFIED reference pipeline code -- topaz.denoise.Denoise.denoise with patches, and the `topaz denoise3d` command
    (Denoise3D -> PatchDataset crops -> model -> MRC) -- on the drop-in modules with the C handle underneath."""
    import test_dropin_reference_cli as T
    with c_engine_under_hook(monkeypatch) as hook:
        T.test_reference_denoise_pipeline_runs_on_dropin_unet(None, monkeypatch)
        n2d = hook.ops.count(0)
        T.test_reference_denoise3d_command_runs_on_dropin_modules(None, tmp_path)
    assert n2d >= 9 and hook.ops.count(0) > n2d    # 3 x 3 patches of the 150 x 170 micrograph, then the tomogram's patches
