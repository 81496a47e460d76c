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

pytestmark = [pytest.mark.gpu,
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


def test_unet3d_c_handle_matches_python_plans_in_fast_precision(c_engine):
    from topaz_b200.denoising.models import UDenoiseNet3D
    c_engine('py', 'fast')
    g = gold('unet3d_seeded')
    m = _load(UDenoiseNet3D(nf=48, base_width=7, top_width=3), seeded_state(unet_shapes(48, 7, 3, 3), int(g['seed'])))
    for x in (torch.from_numpy(g['x']), torch.randn(1, 1, 32, 40, 36, generator=torch.Generator().manual_seed(5))):
        y_c, y_py = _both(m, x.cuda(), c_engine)
        assert torch.equal(y_c, y_py)


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
