"""-m gpu tests of the U-Net entry points of the model-level C ABI (tpz_unet_create / tpz_unet2d_forward / tpz_unet3d_forward):
the denoisers routed through the native handle (TPZ_UNET_ENGINE=c) against the Python-built plans (bit-identical: same kernels,
same packed bytes, same launch order -- tests/test_unet_abi.py proves that on the CPU simulation) and against the reference goldens.

STATUS: these entry points were written after round 2's GPU budget was spent; they have not run on hardware yet, which is why the
C path is opt-in and why the tests below are non-strict xfail (a first failure here must not mask the 150+ validated tests; an XPASS
is the first hardware confirmation)."""
import numpy as np
import pytest
import torch

from common import gold, weights_of, seeded_state, check_parity
from common_shapes import unet_shapes

# collected after every other -m gpu file (alphabetical order), so a failure here cannot disturb the validated tests' CUDA context
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600),
              pytest.mark.xfail(strict=False, reason='tpz_unet*_forward: first run on hardware (written after the GPU budget was spent)')]


def _load(model, sd):
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return model.eval().cuda()


@pytest.fixture
def c_engine():
    from topaz_b200 import engine
    saved = engine.UNET_ENGINE, engine.PRECISION

    def switch(which, precision=None):
        engine.UNET_ENGINE = which
        if precision is not None:
            engine.PRECISION = precision
    yield switch
    engine.UNET_ENGINE, engine.PRECISION = saved


def _both(model, x, switch, stats=None):
    from topaz_b200 import engine
    with torch.no_grad():
        switch('py')
        y_py = engine.unet_forward(model, x, stats)
        switch('c')
        y_c = engine.unet_forward(model, x, stats)
        assert model.__dict__['_tpz_plans']['unet_c'][1] is not None, 'the C handle was not used'
    torch.cuda.synchronize()
    return y_c, y_py


def test_unet2d_c_handle_matches_python_plans_and_goldens(c_engine):
    from topaz_b200.denoising.models import UDenoiseNet, UDenoiseNetSmall
    for name, make in (('unet_pretrained', lambda: UDenoiseNet(base_width=11, top_width=5)),
                       ('unet_small_pretrained', lambda: UDenoiseNetSmall(width=11, top_width=5))):
        g = gold(name)
        m = _load(make(), weights_of(g))
        for xk, yk in (('x', 'y'), ('xo', 'yo')):
            y_c, y_py = _both(m, torch.from_numpy(g[xk]).cuda(), c_engine)
            assert torch.equal(y_c, y_py), (name, xk)
            check_parity(y_c.cpu().numpy(), g[yk], 2e-3, f'{name} {xk} (C handle)')
        stats = torch.tensor([10.0, 3.0], device='cuda')
        y_c, y_py = _both(m, torch.from_numpy(g['xo']).cuda(), c_engine, stats)
        assert torch.equal(y_c, y_py), name


@pytest.mark.parametrize('precision', ['auto', 'fast', 'strict'])
def test_unet3d_c_handle_matches_python_plans(precision, c_engine):
    """auto (the 3-D default: split operands in the last four convolutions), fast and strict"""
    from topaz_b200.denoising.models import UDenoiseNet3D
    c_engine('py', precision)
    g = gold('unet3d_seeded')
    m = _load(UDenoiseNet3D(nf=48, base_width=7, top_width=3), seeded_state(unet_shapes(48, 7, 3, 3), int(g['seed'])))
    for x in (torch.from_numpy(g['x']), torch.randn(1, 1, 32, 40, 36, generator=torch.Generator().manual_seed(5))):
        y_c, y_py = _both(m, x.cuda(), c_engine)
        assert torch.equal(y_c, y_py)        # (parity of the Python plans against the goldens: tests/test_gpu_parity*.py)


def test_unet2d_strict_c_handle_matches_python_plans(c_engine):
    from topaz_b200.denoising.models import UDenoiseNet
    c_engine('py', 'strict')
    g = gold('unet_pretrained')
    m = _load(UDenoiseNet(base_width=11, top_width=5), weights_of(g))
    y_c, y_py = _both(m, torch.from_numpy(g['xo']).cuda(), c_engine)
    assert torch.equal(y_c, y_py)
    check_parity(y_c.cpu().numpy(), g['yo'], 1e-3, 'unet pretrained strict (C handle)')


def test_patched_denoise_through_the_c_handle_with_graph_replay(c_engine):
    """Denoise.denoise with patches: every crop shape is captured into a CUDA graph whose launches come from tpz_unet2d_forward."""
    from topaz_b200.denoise import Denoise
    from topaz_b200.denoising.models import UDenoiseNet
    g = gold('unet_pretrained')
    img = g['img']
    c_engine('py')
    d_py = Denoise(_load(UDenoiseNet(base_width=11, top_width=5), weights_of(g)))
    y_py = d_py.denoise(img, patch_size=64, padding=24)
    c_engine('c')
    d_c = Denoise(_load(UDenoiseNet(base_width=11, top_width=5), weights_of(g)))
    y_c = d_c.denoise(img, patch_size=64, padding=24)
    y_c2 = d_c.denoise(img, patch_size=64, padding=24)          # second pass: graph replays only
    assert d_c.model.__dict__['_tpz_plans']['unet_c'][1] is not None
    assert np.array_equal(y_c, y_py) and np.array_equal(y_c2, y_py)


def _write_unet_file(path, model, x, y_ref, tol):
    """serialise a denoiser + patch + expected output the way tests/c/denoise_c.c reads them"""
    import struct
    from torch import nn
    enc = [getattr(model, f'enc{i}') for i in range(1, 10) if hasattr(model, f'enc{i}')]
    depth = len(enc)
    dims = 3 if isinstance(enc[0][0], nn.Conv3d) else 2
    convs = [e[0] for e in enc]
    for l in range(depth - 1, 0, -1):
        convs += [getattr(model, f'dec{l}')[0], getattr(model, f'dec{l}')[2]]
    convs.append(model.dec1[4])
    out = [struct.pack('<2i', dims, depth)]
    for c in convs:
        w = c.weight.detach().cpu().float()
        out.append(struct.pack('<4i', w.shape[0], w.shape[1], w.shape[-1], int(c.bias is not None)))
        out.append(w.numpy().astype('<f4').tobytes())
        if c.bias is not None:
            out.append(c.bias.detach().cpu().float().numpy().astype('<f4').tobytes())
    x = np.ascontiguousarray(x, dtype='<f4')
    shape = (1, 1) + x.shape if dims == 2 else (1,) + x.shape
    out.append(struct.pack('<4i', *shape)); out.append(struct.pack('<f', tol))
    out.append(x.tobytes()); out.append(np.ascontiguousarray(y_ref, dtype='<f4').tobytes())
    with open(path, 'wb') as fh:
        fh.write(b''.join(out))


def test_plain_c_program_denoises_the_reference_golden(tmp_path):
    """tests/c/denoise_c.c: mean/std, normalise, tpz_unet2d_forward with the de-normalising epilogue, against the reference's own
    Denoise._denoise output (golden y_call) -- no Python, no torch in that process."""
    import os
    import subprocess
    from common import ROOT
    from topaz_b200.denoising.models import UDenoiseNet, UDenoiseNetSmall
    exe = os.path.join(ROOT, 'build', 'denoise_c')
    if not os.path.exists(exe):
        import __graft_entry__ as ge
        ge.build_c_tests()
    for name, make in (('unet_pretrained', lambda: UDenoiseNet(base_width=11, top_width=5)),
                       ('unet_small_pretrained', lambda: UDenoiseNetSmall(width=11, top_width=5))):
        g = gold(name)
        m = make()
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in weights_of(g).items()})
        path = str(tmp_path / f'{name}.bin')
        _write_unet_file(path, m, g['img'], g['y_call'], 1e-3)
        env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, 'topaz_b200') + ':' + os.environ.get('LD_LIBRARY_PATH', ''))
        r = subprocess.run([exe, path], capture_output=True, text=True, env=env, timeout=300)
        print(r.stdout.strip(), r.stderr.strip())
        assert r.returncode == 0, (name, r.returncode, r.stdout, r.stderr)
