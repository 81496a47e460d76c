"""Pin the CPU oracle (oracle/topaz_oracle.py) against goldens produced by the real reference
(tools/make_goldens.py).  fp32 CPU vs fp32 CPU: tolerance 2e-5 relative (op re-association only)."""
import numpy as np
import pytest
import torch

from common import gold, weights_of, seeded_state, rel_err
from oracle import topaz_oracle as O

TOL = 2e-5


def _close(y, ref, tol=TOL):
    m, l2 = rel_err(y, ref)
    assert m <= tol and l2 <= tol, (m, l2)


def test_resnet8_u32_pretrained_dense_and_crops():
    g = gold('resnet8_u32_pretrained'); sd = weights_of(g)
    assert O.resnet_width(O.resnet_spec('resnet8', 32)) == int(g['width']) == 71
    _close(O.classifier_forward(sd, g['x'], 'resnet8', 32, filled=True).numpy(), g['y_dense'])
    _close(O.classifier_forward(sd, g['crops'], 'resnet8', 32, filled=False).numpy(), g['y_crops'])


def test_resnet8_u32_patched_scoring_matches_reference():
    g = gold('resnet8_u32_pretrained'); sd = weights_of(g)
    fwd = lambda p: O.classifier_forward(sd, p, 'resnet8', 32, filled=True)
    y = O.score_in_patches(fwd, torch.from_numpy(g['xp']), 64 + 2 * 35, 35)
    assert y.dtype == np.float64
    _close(y, g['y_patch'])


def test_nms_bit_exact_on_reference_scores():
    g = gold('resnet8_u32_pretrained')
    s, c = O.nms(g['y_full'][0, 0], 6, -6.0)
    assert np.array_equal(c, g['nms_coords']) and np.array_equal(s, g['nms_scores'])


def test_resnet8_u64_pretrained_dense():
    g = gold('resnet8_u64_pretrained')
    _close(O.classifier_forward(weights_of(g), g['x'], 'resnet8', 64, filled=True).numpy(), g['y_dense'])


@pytest.mark.parametrize('name,arch,units,scaling,bn', [
    ('resnet16_u16', 'resnet16', 16, 1, False),
    ('resnet8_u16_bn', 'resnet8', 16, 1, True),
    ('conv31_u16x2', 'conv31', 16, 2, True),
    ('conv63_u32x2', 'conv63', 32, 2, True),
    ('conv63_u16_nobn', 'conv63', 16, 1, False),
])
def test_seeded_classifiers(name, arch, units, scaling, bn):
    g = gold('cls_' + name)
    # shapes are recovered by building the product-independent shape table from the oracle spec
    from common_shapes import classifier_shapes
    shapes = classifier_shapes(arch, units, scaling, bn)
    assert list(shapes.keys()) == [str(k) for k in g['keys']]
    sd = seeded_state(shapes, int(g['seed']))
    _close(O.classifier_forward(sd, g['xc'], arch, units, False, bn, scaling).numpy(), g['yc'])
    _close(O.classifier_forward(sd, g['xd'], arch, units, True, bn, scaling).numpy(), g['yd'])


def test_unet_pretrained():
    g = gold('unet_pretrained'); sd = weights_of(g)
    _close(O.unet_forward(sd, g['x']).numpy(), g['y'])
    _close(O.unet_forward(sd, g['xo']).numpy(), g['yo'])
    _close(O.denoise_call(sd, g['img']), g['y_call'])
    _close(O.denoise(sd, g['img'], patch_size=64, padding=24), g['y_pat'])


def test_unet_seeded_and_3d():
    from common_shapes import unet_shapes
    g = gold('unet_seeded_nf16')
    sh = unet_shapes(16, 7, 3, 2)
    assert list(sh.keys()) == [str(k) for k in g['keys']]
    _close(O.unet_forward(seeded_state(sh, int(g['seed'])), g['x']).numpy(), g['y'])
    g = gold('unet3d_seeded')
    sh = unet_shapes(48, 7, 3, 3)
    assert list(sh.keys()) == [str(k) for k in g['keys']]
    sd = seeded_state(sh, int(g['seed']))
    _close(O.unet_forward(sd, g['x']).numpy(), g['y'])
    _close(O.denoise3d(sd, g['tomo'], 16, 8), g['y_tomo'], 5e-5)


def test_ge_binomial_three_steps():
    g = gold('ge_binomial_u32'); sd = weights_of(gold('resnet8_u32_pretrained'))
    B = int(g['B'])
    Xs = [np.random.default_rng(4000 + s).standard_normal((B, 71, 71)).astype(np.float32) for s in range(3)]
    outs, grads, final = O.ge_binomial_steps(sd, Xs, [g['Y']] * 3, 'resnet8', 32, float(g['pi']))
    np.testing.assert_allclose(np.array(outs), g['outs'], rtol=2e-4, atol=1e-6)
    for k in sd:
        _close(grads[0][k], g['g1.' + k], 2e-4)
        _close(final[k], g['p3.' + k], 1e-5)


def test_ge_binomial_batchnorm_steps_match_reference():
    """Training-mode BatchNorm in the oracle (batch statistics, running buffers) vs the reference's 3 GE_binomial steps of
    ResNet8(units=32, bn=True) -- the default model of `topaz train` (commands/train.py:89-91).  Both run the same fp32
    torch CPU operators, so no ReLU mask flips separate them; the imposed-mask variant used by the GPU gradient test
    must give the same gradient when fed the oracle's own masks."""
    import torch
    from common_shapes import classifier_shapes
    g = gold('ge_binomial_u32_bn')
    sd = seeded_state(classifier_shapes('resnet8', 32, 1, True), int(g['seed']))
    B = int(g['B'])
    Xs = [np.random.default_rng(4000 + s).standard_normal((B, 71, 71)).astype(np.float32) for s in range(3)]
    outs, grads, final = O.ge_binomial_steps(sd, Xs, [g['Y']] * 3, 'resnet8', 32, float(g['pi']), bn=True)
    np.testing.assert_allclose(np.array(outs), g['outs'], rtol=2e-4, atol=1e-6)
    for k in grads[0]:
        _close(grads[0][k], g['g1.' + k], 5e-4)
    for k in final:
        if k.endswith('num_batches_tracked'):
            assert int(final[k]) == 3
        else:
            _close(final[k], g['p3.' + k], 2e-4)
    # imposed masks == derived masks -> identical gradient
    params = {k: torch.from_numpy(v).clone().requires_grad_(v.dtype == np.float32 and 'running' not in k) for k, v in sd.items()}
    masks = []
    with torch.no_grad():
        x = torch.from_numpy(Xs[0]).unsqueeze(1)
        spec = O.resnet_spec('resnet8', 32)
        # derive the masks by running the oracle with a recording activation
        rec = []
        orig = O._act
        O._act = lambda y, rm: (rec.append((y > 0)), orig(y, None))[1]
        try:
            O.classifier_forward_grad(params, torch.from_numpy(Xs[0]), 'resnet8', 32, bn=True)
        finally:
            O._act = orig
        masks = rec
    assert len(masks) == 8        # conv, 3 x (act0, act1), conv
    score = O.classifier_forward_grad(params, torch.from_numpy(Xs[0]), 'resnet8', 32, bn=True, relu_masks=masks).view(-1)
    _, _, loss = O.ge_binomial_loss(score, torch.from_numpy(g['Y']), float(g['pi']), 1.0)
    loss.backward()
    for k in grads[0]:
        _close(params[k].grad.numpy(), grads[0][k], 1e-6)


@pytest.mark.parametrize('tag,bn', [('ge_binomial_conv31_bn', True), ('ge_binomial_conv31_nobn', False)])
def test_ge_binomial_prelu_extractor_steps_match_reference(tag, bn):
    """conv31 (conv -> [BN] -> PReLU, learnable slope) in training: the oracle's two GE_binomial steps vs the reference's."""
    from common_shapes import classifier_shapes
    g = gold(tag)
    sd = seeded_state(classifier_shapes('conv31', 16, 2, bn), int(g['seed']))
    assert list(sd.keys()) == [str(k) for k in g['keys']]
    B, W = int(g['B']), int(g['width'])
    Xs = [np.random.default_rng(4200 + s).standard_normal((B, W, W)).astype(np.float32) for s in range(2)]
    outs, grads, final = O.ge_binomial_steps(sd, Xs, [g['Y']] * 2, 'conv31', 16, float(g['pi']), bn=bn, unit_scaling=2)
    np.testing.assert_allclose(np.array(outs), g['outs'], rtol=2e-4, atol=1e-6)
    for k in grads[0]:
        _close(grads[0][k], g['g1.' + k], 5e-4)
    for k in final:
        if not k.endswith('num_batches_tracked'):
            _close(final[k], g['p2.' + k], 2e-4)


def test_dropout_training_step_matches_reference_with_its_masks():
    """`topaz train --dropout 0.25`: ResNet8 (16 units, BatchNorm) in train() mode.  With the keep-masks that the reference's
    three nn.Dropout layers drew (recorded by hooks in tools/make_goldens_bn.py) imposed, the oracle reproduces the
    reference's loss tuple and every gradient -- this pins where the Dropout layers sit (after blocks 0, 2, 4), how they shift
    the state_dict keys, and the 1/(1-p) scaling.  The arithmetic itself is also checked against F.dropout."""
    import torch
    import torch.nn.functional as F
    from common import dropout_masks_of
    g = gold('ge_binomial_u16_dropout')
    p = float(g['p'])
    keys = [str(k) for k in g['keys']]
    assert 'features.features.2.conv0.weight' in keys and 'features.features.1.conv0.weight' not in keys   # slot 1 = Dropout
    shapes = {k: tuple(g['g1.' + k].shape) for k in keys if 'g1.' + k in g.files}
    # buffers have no gradient entry: rebuild the full shape table from a BN ResNet8 with shifted indices
    from common_shapes import classifier_shapes
    base = classifier_shapes('resnet8', 16, 1, True)
    remap = {0: 0, 1: 2, 2: 3, 3: 5, 4: 6}
    full = {}
    for k, shp in base.items():
        if k.startswith('features.features.'):
            parts = k.split('.')
            parts[2] = str(remap[int(parts[2])])
            k = '.'.join(parts)
        full[k] = shp
    assert list(full.keys()) == keys
    sd = seeded_state(full, int(g['seed']))
    B = int(g['B'])
    X = np.random.default_rng(4300).standard_normal((B, 71, 71)).astype(np.float32)
    params = {k: torch.from_numpy(v).clone().requires_grad_('running' not in k and v.dtype == np.float32) for k, v in sd.items()}
    masks = [torch.from_numpy(m) for m in dropout_masks_of(g)]
    score = O.classifier_forward_grad(params, torch.from_numpy(X), 'resnet8', 16, bn=True, dropout=p, dropout_masks=masks).view(-1)
    Y = torch.from_numpy(g['Y'])
    cls, ge, loss = O.ge_binomial_loss(score, Y, float(g['pi']), 1.0)
    loss.backward()
    prec, tpr, fpr = O.ge_binomial_metrics(score.detach(), Y)
    np.testing.assert_allclose([cls.item(), ge.item(), prec, tpr, fpr], g['out'], rtol=2e-4, atol=1e-6)
    for k in shapes:
        _close(params[k].grad.numpy(), g['g1.' + k], 5e-4)
    x = torch.randn(4, 3, 5, 5)
    torch.manual_seed(3)
    ref = F.dropout(x, p, training=True)
    assert torch.equal(O._dropout(x, p, iter([ref != 0])), ref)


def test_filters():
    g = gold('filters')
    _close(O.gaussian_denoise(g['img'], 1.5), g['gauss'])
    n, mu, std = O.affine_normalize(g['img'])
    _close(n, g['norm']); assert abs(mu - float(g['mu'])) < 1e-6 and abs(std - float(g['std'])) < 1e-6


def test_log_binom_matches_scipy():
    scipy_stats = pytest.importorskip('scipy.stats')
    for N, pi in [(240, 0.035), (60, 0.2), (1, 0.5)]:
        ref = scipy_stats.binom.logpmf(np.arange(N + 1), N, pi).astype(np.float32)
        np.testing.assert_allclose(O.log_binom_pmf(N, pi), ref, rtol=2e-6, atol=1e-5)


def test_nms3d_bit_exact_vs_reference_golden():
    g = gold('nms3d')
    for tag in 'abcd':
        s, c = O.nms3d(g[f'{tag}.x'], float(g[f'{tag}.r']), float(g[f'{tag}.scale']), float(g[f'{tag}.thr']))
        assert len(s) > 0 and np.array_equal(c, g[f'{tag}.coords']) and np.array_equal(s, g[f'{tag}.scores']), tag


def test_downsample_and_gmm_normalize_vs_reference_golden():
    g = gold('preprocess')
    x = g['x']
    for tag, kw in [('f2', dict(factor=2)), ('f4', dict(factor=4)), ('f3', dict(factor=3)), ('s', dict(shape=(37, 50))),
                    ('f1p7', dict(factor=1.7))]:
        y = O.downsample(x, **kw)
        assert y.shape == g['ds.' + tag].shape and np.abs(y - g['ds.' + tag]).max() < 1e-4 * np.abs(g['ds.' + tag]).max()
    for tag in 'abc':
        y, mu, std, pi = O.gmm_normalize(g[f'n.{tag}.x'], float(g[f'n.{tag}.alpha']), float(g[f'n.{tag}.beta']), int(g[f'n.{tag}.iters']))
        assert abs(mu - g[f'n.{tag}.mu']) < 1e-5 * abs(mu) and abs(std - g[f'n.{tag}.std']) < 1e-4 * std, tag
        assert abs(pi - g[f'n.{tag}.pi']) < 1e-4
        assert np.abs(y - g[f'n.{tag}.y']).max() < 1e-3


# ---------------------------------------------------------------- round-2 fixtures (tools/make_goldens_r2.py)
def test_resnet16_u64_pretrained():
    """`topaz extract`'s default model (commands/extract.py:18 -> factory.py:34-36) with its packaged weights."""
    g = gold('resnet16_u64_pretrained'); sd = weights_of(g)
    assert O.resnet_width(O.resnet_spec('resnet16', 64)) == int(g['width'])
    _close(O.classifier_forward(sd, g['x'], 'resnet16', 64, filled=True).numpy(), g['y_dense'])
    _close(O.classifier_forward(sd, g['crops'], 'resnet16', 64, filled=False).numpy(), g['y_crops'])


def test_conv127_seeded():
    from common_shapes import classifier_shapes
    g = gold('cls_conv127_u16x2')
    shapes = classifier_shapes('conv127', 16, 2, True)
    assert list(shapes.keys()) == [str(k) for k in g['keys']]
    sd = seeded_state(shapes, int(g['seed']))
    _close(O.classifier_forward(sd, g['xc'], 'conv127', 16, False, True, 2).numpy(), g['yc'])
    _close(O.classifier_forward(sd, g['xd'], 'conv127', 16, True, True, 2).numpy(), g['yd'])


def test_unet_small_pretrained():
    """UDenoiseNetSmall (denoising/models.py:178-244) with the packaged `unet-small` weights."""
    g = gold('unet_small_pretrained'); sd = weights_of(g)
    _close(O.unet_forward(sd, g['x']).numpy(), g['y'])
    _close(O.unet_forward(sd, g['xo']).numpy(), g['yo'])
    _close(O.denoise_call(sd, g['img']), g['y_call'])
    _close(O.denoise(sd, g['img'], patch_size=64, padding=24), g['y_pat'])


def test_unet3d_pretrained_10a_small_block_and_tomogram():
    """UDenoiseNet3D with the packaged `unet-3d-10a` weights (denoising/models.py:452-564, 576-578): a 32^3 block and
    Denoise3D.denoise on a 70x50x64 tomogram.  (The 192^3 patch of the fixture is checked against the CUDA path on the GPU;
    the oracle needs ~20 s for it, see test_unet3d_pretrained_192_patch_oracle_windows.)"""
    g = gold('unet3d_pretrained_10a'); sd = weights_of(g)
    _close(O.unet_forward(sd, g['x32']).numpy(), g['y32'])
    _close(O.denoise3d(sd, g['tomo'], 32, 16), g['y_tomo'], 5e-5)
