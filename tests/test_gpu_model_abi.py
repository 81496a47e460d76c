"""-m gpu tests of the model-level C ABI (include/topaz_b200.h: tpz_model_create / _update_weights / _destroy,
tpz_workspace_bytes, tpz_resnet_dense_forward):
  * a plain-C host program (tests/c/score_c.c, built by __graft_entry__.build(); no Python, no torch) scores the reference's
    golden image and a 512 x 512 micrograph (cfg1) with the pretrained resnet8_u64 to 1e-3
  * the Python engine routed through the same handle gives bit-identical scores to its own Python-built plans, for ResNet8/16,
    BatchNorm models and the PReLU conv extractors; the on-device repack follows parameter updates."""
import os
import struct
import subprocess

import numpy as np
import pytest
import torch

from common import ROOT, gold, weights_of, seeded_state, check_parity
from common_shapes import classifier_shapes
from oracle import topaz_oracle as O

pytestmark = pytest.mark.gpu


def _classifier(arch, units, scaling=1, bn=False):
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    kw = dict(units=units, bn=bn)
    if arch.startswith('conv'):
        kw['unit_scaling'] = scaling
    return LinearClassifier(get_feature_extractor(arch, **kw))


def _load(model, sd):
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return model


def _write_model_file(path, model, x, y_ref):
    """serialise a FILLED model the way tests/c/score_c.c reads it"""
    from topaz_b200 import engine
    blocks = engine._feature_blocks(model.features)
    f32 = lambda t: t.detach().cpu().float().numpy().astype('<f4').tobytes()
    bn4 = lambda bn: b''.join(f32(t) for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var))
    out = [struct.pack('<i', len(blocks))]
    for b in blocks:
        if b['kind'] == 'conv':
            w = b['w']
            out.append(struct.pack('<11i', 0, w.shape[1], w.shape[0], w.shape[-1], b['dil'], 1, int(b['b'] is not None), 0, 0,
                                   int(b['bn'] is not None), 0))
            out.append(struct.pack('<4f', b['slope'], 0.0, b['bn'].eps if b['bn'] is not None else 0.0, 0.0))
            out.append(f32(w))
            if b['b'] is not None:
                out.append(f32(b['b']))
            if b['bn'] is not None:
                out.append(bn4(b['bn']))
        else:
            w0, w1 = b['w0'], b['w1']
            out.append(struct.pack('<11i', 1, w0.shape[0], w1.shape[0], 3, b['d0'], b['d1'], int(b['b0'] is not None),
                                   int(b['b1'] is not None), int(b['proj'] is not None), int(b['bn0'] is not None), int(b['bn1'] is not None)))
            out.append(struct.pack('<4f', b['slope0'], b['slope1'], b['bn0'].eps if b['bn0'] is not None else 0.0,
                                   b['bn1'].eps if b['bn1'] is not None else 0.0))
            out.append(f32(w0))
            if b['b0'] is not None:
                out.append(f32(b['b0']))
            out.append(f32(w1))
            if b['b1'] is not None:
                out.append(f32(b['b1']))
            if b['proj'] is not None:
                out.append(f32(b['proj']))
            if b['bn0'] is not None:
                out.append(bn4(b['bn0']))
            if b['bn1'] is not None:
                out.append(bn4(b['bn1']))
    cw = model.classifier.weight.detach().reshape(-1)
    out.append(struct.pack('<i', cw.numel())); out.append(f32(cw)); out.append(f32(model.classifier.bias.detach().reshape(-1)))
    B, _, H, W = x.shape
    out.append(struct.pack('<4i', model.features.width // 2, B, H, W))
    out.append(np.ascontiguousarray(x, dtype='<f4').tobytes()); out.append(np.ascontiguousarray(y_ref, dtype='<f4').tobytes())
    with open(path, 'wb') as fh:
        fh.write(b''.join(out))


def test_plain_c_program_scores_golden_and_512_image(tmp_path):
    exe = os.path.join(ROOT, 'build', 'score_c')
    if not os.path.exists(exe):
        import __graft_entry__ as ge
        ge.build_c_tests()
    g = gold('resnet8_u64_pretrained'); sd = weights_of(g)
    m = _load(_classifier('resnet8', 64), sd); m.eval(); m.fill()
    cases = [('golden', g['x'], g['y_dense'])]
    x512 = np.random.default_rng(1234).standard_normal((1, 1, 512, 512)).astype(np.float32)       # cfg1 input
    cases.append(('cfg1_512', x512, O.classifier_forward(sd, x512, 'resnet8', 64, filled=True).numpy()))
    for name, x, ref in cases:
        path = str(tmp_path / f'{name}.bin')
        _write_model_file(path, m, x, ref)
        env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, 'topaz_b200') + ':' + os.environ.get('LD_LIBRARY_PATH', ''))
        r = subprocess.run([exe, path], capture_output=True, text=True, env=env, timeout=300)
        print(r.stdout.strip(), r.stderr.strip())
        assert r.returncode == 0, (name, r.returncode, r.stdout, r.stderr)


@pytest.mark.parametrize('name,arch,units,scaling,bn', [
    ('resnet8_u64_pretrained', 'resnet8', 64, 1, False),
    ('cls_resnet16_u16', 'resnet16', 16, 1, False),
    ('cls_resnet8_u16_bn', 'resnet8', 16, 1, True),
    ('cls_conv63_u32x2', 'conv63', 32, 2, True),
    ('cls_conv127_u16x2', 'conv127', 16, 2, True),
])
def test_c_model_path_is_bit_identical_to_python_plans(name, arch, units, scaling, bn):
    from topaz_b200 import engine
    g = gold(name)
    sd = weights_of(g) if 'pretrained' in name else seeded_state(classifier_shapes(arch, units, scaling, bn), int(g['seed']))
    x = torch.from_numpy(g['x'] if 'pretrained' in name else g['xd']).cuda()
    m = _load(_classifier(arch, units, scaling, bn), sd).cuda(); m.eval(); m.fill()
    old = engine.DENSE_ENGINE
    try:
        engine.DENSE_ENGINE = 'py'
        with torch.no_grad():
            y_py = m(x).cpu()
        engine.DENSE_ENGINE = 'c'
        with torch.no_grad():
            y_c = m(x).cpu()
        print(name, 'python plans vs C model: max |diff|', float((y_py - y_c).abs().max()), 'of', float(y_py.abs().max()))
        assert torch.equal(y_py, y_c), float((y_py - y_c).abs().max())
        # parameters change in place (an optimizer epoch): the handle repacks on the device and follows
        with torch.no_grad():
            for p in m.parameters():
                p.mul_(1.01)
            y_c2 = m(x).cpu()
            engine.DENSE_ENGINE = 'py'
            y_py2 = m(x).cpu()
        assert not torch.equal(y_c2, y_c) and torch.equal(y_py2, y_c2)
    finally:
        engine.DENSE_ENGINE = old


@pytest.mark.parametrize('name,arch,units,scaling,bn', [
    ('cls_resnet8_u16_bn', 'resnet8', 16, 1, True),
    ('cls_conv63_u32x2', 'conv63', 32, 2, True),
    ('resnet8_u64_pretrained', 'resnet8', 64, 1, False),
])
def test_c_model_packs_the_same_bytes_as_the_python_packer(name, arch, units, scaling, bn):
    """Step by step: the fp16 k-block weights and fp32 biases the library packs ON the device (BatchNorm folded by
    bn_affine_kernel / bias_fold_kernel) against the Python packer's (ops.pack_tc_conv on the host)."""
    from topaz_b200 import engine
    from topaz_b200.model_abi import DenseModel
    g = gold(name)
    sd = weights_of(g) if 'pretrained' in name else seeded_state(classifier_shapes(arch, units, scaling, bn), int(g['seed']))
    m = _load(_classifier(arch, units, scaling, bn), sd).cuda(); m.eval(); m.fill()
    plan = engine._build_dense_plan(m.features, m.classifier, torch.device('cuda'))
    dm = DenseModel(m)
    steps = plan['steps']
    assert steps[0]['op'] == 'first_tc'
    w, b = dm.step_buffers(-1)
    dw = (w.float() - steps[0]['w'].float()).abs().max().item()
    db = (b - steps[0]['b']).abs().max().item()
    print(f'first layer: weights max diff {dw:.3e}, bias max diff {db:.3e}')
    worst_w, worst_b = dw, db
    for i, st in enumerate(steps[1:]):
        p = st['plan']
        w, b = dm.step_buffers(i)
        assert tuple(w.shape) == tuple(p.weights.shape), (i, tuple(w.shape), tuple(p.weights.shape))
        dw = (w.float() - p.weights.float()).abs().max().item()
        db = (b - p.bias).abs().max().item()
        nz = int((w != p.weights).sum())
        print(f'step {i}: {tuple(w.shape)} weights max diff {dw:.3e} ({nz} fp16 values differ), bias max diff {db:.3e}')
        worst_w, worst_b = max(worst_w, dw), max(worst_b, db)
    # the launch arguments too: k-block tables, lattice, tap grids, origins, slopes (pointers and the per-forward geometry aside)
    import ctypes as C
    from topaz_b200 import _lib, ops
    x = torch.from_numpy(g['x'] if 'pretrained' in name else g['xd']).cuda()
    with torch.no_grad():
        dm.forward(x[:, 0].contiguous())
    for i, st in enumerate(steps[1:]):
        pa = ops._static_tc_args(st['plan'])
        ca = _lib.TpzTcConvArgs()
        _lib.check(_lib.lib().tpz_model_step_args(dm.handle, i, C.byref(ca)))
        for f in ('nsrc', 'KC', 'nkb', 'Co', 'TW', 'TH', 'lattice', 'phase_sel', 'lattice_z', 'phase_z', 'neg_slope', 'out_lo'):
            assert getattr(pa, f) == getattr(ca, f), (i, f, getattr(pa, f), getattr(ca, f))
        # the fused classifier bias is a launch-time field on the Python side (ops.make_tc_args copies plan.dot_b)
        assert C.c_float(st['plan'].dot_b).value == ca.dot_b, (i, 'dot_b', st['plan'].dot_b, ca.dot_b)
        for j in range(pa.nkb):
            a, b = pa.kb[j], ca.kb[j]
            assert (a.dx, a.dy, a.dz, a.c0, a.src) == (b.dx, b.dy, b.dz, b.c0, b.src), (i, j)
        for sidx in range(pa.nsrc):
            a, b = pa.src[sidx], ca.src[sidx]
            assert (a.C, tuple(a.org), a.kw, a.kh, a.lat, a.no_phase, a.lat_z) == (b.C, tuple(b.org), b.kw, b.kh, b.lat, b.no_phase, b.lat_z), (i, sidx)
    dm.close()
    assert worst_w == 0.0 and worst_b == 0.0, (worst_w, worst_b)


def test_c_model_declines_rows_outside_the_fp16_range():
    """A BatchNorm-folded weight row below 2^-10 (tiny gamma) or beyond 2^14 needs the per-row scale of the Python packer
    (ops._row_scales): the C packer reports it (TPZ_E_WEIGHT_RANGE), the engine falls back to the Python plans for that parameter
    state, and the scores stay right."""
    from topaz_b200 import engine
    g = gold('cls_resnet8_u16_bn')
    sd = seeded_state(classifier_shapes('resnet8', 16, 1, True), int(g['seed']))
    sd['features.features.1.bn0.weight'] = sd['features.features.1.bn0.weight'] * np.float32(1e-6)
    m = _load(_classifier('resnet8', 16, 1, True), sd).cuda(); m.eval(); m.fill()
    x = torch.from_numpy(g['xd']).cuda()
    old = engine.DENSE_ENGINE
    try:
        engine.DENSE_ENGINE = 'c'
        with torch.no_grad():
            y_c = m(x).cpu()
        assert m.__dict__['_tpz_plans']['dense_c'][2] is None          # declined, remembered
        engine.DENSE_ENGINE = 'py'
        with torch.no_grad():
            y_py = m(x).cpu()
    finally:
        engine.DENSE_ENGINE = old
    assert torch.isfinite(y_c).all() and torch.equal(y_c, y_py)
    ref = O.classifier_forward(sd, g['xd'], 'resnet8', 16, filled=True, bn=True).numpy()
    check_parity(y_c.numpy(), ref, 3e-3, 'BN gamma x1e-6 (row-scaled plans)')
