"""not-gpu tests (kernel launches replaced by tests/sim_backend.py) of the round-2 parity features:
  * strict mode (engine.STRICT): (hi, lo) fp16 operand pairs -> every seeded network meets the north-star 1e-3
  * fp16 range guard: un-normalised inputs (|x| ~ 1e5) and extreme BatchNorm-folded weights are scored like the fp32 reference
  * the round-2 reference goldens: conv127 (dilation 16), pretrained resnet16_u64, unet-small, unet-3d-10a."""
import numpy as np
import pytest
import torch

from common import gold, weights_of, seeded_state, rel_err, check_parity
from common_shapes import classifier_shapes, unet_shapes
from oracle import topaz_oracle as O
import sim_backend

TOL = 1e-3


def _load(model, sd):
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return model


def _classifier(arch, units, scaling=1, bn=False):
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    kw = dict(units=units, bn=bn)
    if arch.startswith('conv'):
        kw['unit_scaling'] = scaling
    return LinearClassifier(get_feature_extractor(arch, **kw))


@pytest.fixture
def strict_mode():
    from topaz_b200 import engine
    old = engine.PRECISION
    engine.PRECISION = 'strict'
    yield
    engine.PRECISION = old


@pytest.mark.parametrize('name,arch,units,scaling,bn', [
    ('resnet16_u16', 'resnet16', 16, 1, False),
    ('resnet8_u16_bn', 'resnet8', 16, 1, True),
    ('conv31_u16x2', 'conv31', 16, 2, True),
    ('conv63_u32x2', 'conv63', 32, 2, True),
    ('conv127_u16x2', 'conv127', 16, 2, True),
])
def test_strict_mode_seeded_classifiers_meet_1e3(strict_mode, name, arch, units, scaling, bn):
    g = gold('cls_' + name)
    m = _load(_classifier(arch, units, scaling, bn), seeded_state(classifier_shapes(arch, units, scaling, bn), int(g['seed'])))
    m.eval(); m.fill()
    with sim_backend.patched(), torch.no_grad():
        y = m(torch.from_numpy(g['xd'])).numpy()
    mx, l2 = check_parity(y, g['yd'], TOL, 'strict ' + name)
    assert mx < 1e-4 and l2 < 1e-4          # 22-bit operands: two orders below the default mode


def test_strict_mode_unets(strict_mode):
    from topaz_b200.denoising.models import UDenoiseNet, UDenoiseNet3D, UDenoiseNetSmall, DenoiseNet2
    g = gold('unet_seeded_nf16')
    m = _load(UDenoiseNet(nf=16, base_width=7, top_width=3), seeded_state(unet_shapes(16, 7, 3, 2), int(g['seed']))); m.eval()
    with sim_backend.patched(), torch.no_grad():
        y = m(torch.from_numpy(g['x'])).numpy()
    check_parity(y, g['y'], 1e-4, 'strict unet nf16')
    g = gold('unet_pretrained')
    m = _load(UDenoiseNet(base_width=11, top_width=5), weights_of(g)); m.eval()
    with sim_backend.patched(), torch.no_grad():
        yo = m(torch.from_numpy(g['xo'])).numpy()           # odd sizes: materialised up-sampling path
        y = m(torch.from_numpy(g['x'][:, :, :64, :64].copy())).numpy()
    check_parity(yo, g['yo'], 1e-4, 'strict unet pretrained odd')
    check_parity(y, O.unet_forward(weights_of(g), g['x'][:, :, :64, :64]).numpy(), 1e-4, 'strict unet pretrained 64x64')
    g = gold('unet3d_seeded')
    m = _load(UDenoiseNet3D(nf=48, base_width=7, top_width=3), seeded_state(unet_shapes(48, 7, 3, 3), int(g['seed']))); m.eval()
    with sim_backend.patched(), torch.no_grad():
        y = m(torch.from_numpy(g['x'])).numpy()
    check_parity(y, g['y'], 1e-4, 'strict unet3d seeded')
    g = gold('fcnn_affine_seeded')
    mf = _load(DenoiseNet2(64, width=11), seeded_state({k: tuple(v.shape) for k, v in DenoiseNet2(64, width=11).state_dict().items()}, 301)); mf.eval()
    with sim_backend.patched(), torch.no_grad():
        y = mf(torch.from_numpy(g['x'])).numpy()
    check_parity(y, g['y_fcnn'], 1e-4, 'strict fcnn')


def test_conv127_dilation16_dense_sim():
    """conv127 filled: dilations 1,2,4,8,16 -- the last layer is outside the halo-resident kernel's lattice range and takes
    the per-tap kernel on the GPU; here the plan / k-block tables are checked against the reference golden."""
    g = gold('cls_conv127_u16x2')
    m = _classifier('conv127', 16, 2, True)
    assert list(m.state_dict().keys()) == [str(k) for k in g['keys']]
    _load(m, seeded_state(classifier_shapes('conv127', 16, 2, True), int(g['seed']))); m.eval()
    assert m.width == int(g['width']) == 127 and m.fill() == int(g['fill_stride']) == 16
    assert max(c.dilation[0] for c in m.features.features if isinstance(c, torch.nn.Conv2d)) == 16
    with sim_backend.patched(), torch.no_grad():
        y = m(torch.from_numpy(g['xd'])).numpy()
    check_parity(y, g['yd'], 3e-3, 'conv127 dense')


def test_resnet16_u64_pretrained_sim_and_oracle():
    g = gold('resnet16_u64_pretrained'); sd = weights_of(g)
    ref = O.classifier_forward(sd, g['x'], 'resnet16', 64, filled=True).numpy()
    check_parity(ref, g['y_dense'], 2e-5, 'oracle resnet16_u64')
    m = _load(_classifier('resnet16', 64), sd); m.eval()
    assert m.width == int(g['width']) and m.fill() == int(g['fill_stride'])
    with sim_backend.patched(), torch.no_grad():
        y = m(torch.from_numpy(g['x'])).numpy()
    check_parity(y, g['y_dense'], TOL, 'resnet16_u64 dense')


def test_unet_small_and_unet3d_pretrained_sim():
    from topaz_b200.denoising.models import UDenoiseNetSmall, UDenoiseNet3D
    g = gold('unet_small_pretrained'); sd = weights_of(g)
    m = _load(UDenoiseNetSmall(width=11, top_width=5), sd); m.eval()
    with sim_backend.patched(), torch.no_grad():
        y = m(torch.from_numpy(g['x'])).numpy(); yo = m(torch.from_numpy(g['xo'])).numpy()
    check_parity(y, g['y'], TOL, 'unet-small'); check_parity(yo, g['yo'], TOL, 'unet-small odd')
    g = gold('unet3d_pretrained_10a'); sd = weights_of(g)
    m = _load(UDenoiseNet3D(base_width=7), sd); m.eval()
    from topaz_b200 import engine
    assert engine.PRECISION == 'auto'
    with sim_backend.patched(), torch.no_grad():
        y = m(torch.from_numpy(g['x32'])).numpy()
    # default ('auto') precision: split operands in the last four convolutions of the 3-D U-Net (engine._unet_precision)
    check_parity(y, g['y32'], TOL, 'unet-3d-10a 32^3 (auto precision)')
    engine.PRECISION = 'fast'
    try:
        with sim_backend.patched(), torch.no_grad():
            yf = m(torch.from_numpy(g['x32'])).numpy()
    finally:
        engine.PRECISION = 'auto'
    mx, l2 = rel_err(yf, g['y32'])
    print(f'unet-3d-10a 32^3 in fast (11-bit operand) mode: max-rel {mx:.2e} rel-L2 {l2:.2e}  <- why auto mode exists')
    assert 1e-3 < mx < 1e-2


def test_range_guard_large_and_tiny_inputs_sim():
    """An un-normalised micrograph (|x| ~ 1e5 >> fp16 max) and a tiny-scale one: the reference's fp32 path scores both; the
    fp16 path stores activations multiplied by a power of two (range_scale) so it does too -- same relative accuracy as at
    unit scale."""
    g = gold('resnet8_u32_pretrained'); sd = weights_of(g)
    m = _load(_classifier('resnet8', 32), sd); m.eval(); m.fill()
    for scale in (1e5, 3e-6, 1.0):
        x = (g['x'] * np.float32(scale)).astype(np.float32)
        ref = O.classifier_forward(sd, x, 'resnet8', 32, filled=True).numpy()
        with sim_backend.patched(), torch.no_grad():
            y = m(torch.from_numpy(x)).numpy()
        assert np.isfinite(y).all()
        check_parity(y, ref, TOL, f'range guard x{scale:g}')
    from topaz_b200.denoising.models import UDenoiseNet
    gu = gold('unet_pretrained'); sdu = weights_of(gu)
    mu = _load(UDenoiseNet(base_width=11, top_width=5), sdu); mu.eval()
    x = (gu['x'][:, :, :64, :64] * np.float32(2e5)).astype(np.float32)
    with sim_backend.patched(), torch.no_grad():
        y = mu(torch.from_numpy(x)).numpy()
    check_parity(y, O.unet_forward(sdu, x).numpy(), 2e-3, 'range guard unet x2e5')


def test_bn_running_var_1e12_and_weight_range_checks():
    """A BatchNorm layer with running_var = 1e-12 folds into weight rows 316x larger (gamma / sqrt(var + eps), eps = 1e-5) and
    makes every later activation 316x larger: still inside the fp16 range, and scored like the fp32 reference.  Weight rows
    whose fp16 image would overflow / go subnormal are divided by a power of two that the fp32 epilogue restores
    (TpzTcConvArgs.oscale); weights that cannot be represented at all raise instead of producing inf scores."""
    from topaz_b200 import ops
    g = gold('cls_resnet8_u16_bn')
    sd = seeded_state(classifier_shapes('resnet8', 16, 1, True), int(g['seed']))
    sd['features.features.1.bn1.running_var'][:] = 1e-12
    m = _load(_classifier('resnet8', 16, 1, True), sd); m.eval(); m.fill()
    x = g['xd']
    ref = O.classifier_forward(sd, x, 'resnet8', 16, filled=True, bn=True).numpy()
    with sim_backend.patched(), torch.no_grad():
        y = m(torch.from_numpy(x)).numpy()
    assert np.isfinite(y).all()
    check_parity(y, ref, 3e-3, 'BN running_var=1e-12')
    # row scaling: one conv with rows at 1e6 and at 1e-9 against an fp32 convolution of the same fp16-rounded operands
    gen = torch.Generator().manual_seed(5)
    w = torch.randn(32, 32, 3, 3, generator=gen) * 0.1
    w[3] *= 1e6; w[7] *= 1e-9
    a = torch.randn(1, 1, 12, 12, 32, generator=gen).half()
    ref = torch.nn.functional.conv2d(a[0, 0].permute(2, 0, 1)[None].float(), w)[0].permute(1, 2, 0)
    for c in (0, 3, 7):          # read each channel through the fp32 "dot" epilogue (as an fp16 ACTIVATION 1e6 / 1e-9 would not fit)
        onehot = torch.zeros(32); onehot[c] = 1.0
        plan = ops.pack_tc_conv([ops.ConvPart(w, 32, 1)], torch.zeros(32), 32, 1.0, 'cpu', dot_w=onehot)
        assert plan.oscale is not None and float(plan.oscale[3]) >= 2.0 ** 14 and float(plan.oscale[7]) <= 2.0 ** -20
        assert float(plan.oscale[0]) == 1.0 and bool(torch.isfinite(plan.weights.float()).all())
        out = torch.zeros((1, 1, 10, 10), dtype=torch.float32)
        sim_backend.tc_conv(plan, [a], (1, 1, 10, 10), out=None, dot_out=out)
        mx, _ = rel_err(out[0, 0].numpy(), ref[:, :, c].numpy())
        assert mx < 2e-3, (c, mx)
    with pytest.raises(RuntimeError):
        ops.pack_tc_conv([ops.ConvPart(torch.full((16, 32, 3, 3), float('inf')), 32, 1)], None, 32, 0.0, 'cpu')
