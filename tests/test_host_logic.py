"""not-gpu tests of the host side: the C-ABI library loads and exports every declared symbol, and the
Python engine (weight packing, k-block tables, BN folding, U-Net wiring, patch logic) reproduces the
reference goldens when its kernel launches are replaced by the CPU simulation in tests/sim_backend.py."""
import os
import re

import numpy as np
import pytest
import torch

from common import ROOT, gold, weights_of, seeded_state, rel_err, assert_params_after_adam
from common_shapes import classifier_shapes, unet_shapes
import sim_backend

TOL = 1e-3   # fp16 operand rounding is simulated, so the north-star tolerance applies (pretrained weights)
TOL_SEEDED = 3e-3   # He-normal random weights amplify the 2^-11 operand rounding through 9-16 layers


def _load(model, sd):
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return model


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from topaz_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'topaz_b200.h')).read()
    declared = set(re.findall(r'\b(tpz_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations found'
    l = _lib.lib()
    for name in declared:
        assert hasattr(l, name), f'{name} declared in include/topaz_b200.h but not exported'
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)


def test_cpu_tensor_is_rejected():
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    m = LinearClassifier(get_feature_extractor('resnet8', units=16, bn=False)); m.eval(); m.fill()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 80, 80))


def _classifier(arch, units, scaling, bn):
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    kw = dict(units=units, bn=bn)
    if arch.startswith('conv'):
        kw['unit_scaling'] = scaling
    return LinearClassifier(get_feature_extractor(arch, **kw))


def test_dense_resnet8_pretrained_sim():
    g = gold('resnet8_u32_pretrained')
    m = _load(_classifier('resnet8', 32, 1, False), weights_of(g)); m.eval()
    assert m.width == int(g['width'])
    assert m.fill() == 4
    with sim_backend.patched(), torch.no_grad():
        y = m(torch.from_numpy(g['x'])).numpy()
    mx, l2 = rel_err(y, g['y_dense'])
    assert y.shape == g['y_dense'].shape and mx < TOL and l2 < TOL, (mx, l2)


@pytest.mark.parametrize('name,arch,units,scaling,bn', [
    ('resnet16_u16', 'resnet16', 16, 1, False),
    ('resnet8_u16_bn', 'resnet8', 16, 1, True),
    ('conv31_u16x2', 'conv31', 16, 2, True),
    ('conv63_u32x2', 'conv63', 32, 2, True),
    ('conv63_u16_nobn', 'conv63', 16, 1, False),
])
def test_dense_seeded_classifiers_sim(name, arch, units, scaling, bn):
    g = gold('cls_' + name)
    m = _classifier(arch, units, scaling, bn)
    assert [k for k in m.state_dict().keys()] == [str(k) for k in g['keys']]
    _load(m, seeded_state(classifier_shapes(arch, units, scaling, bn), int(g['seed']))); m.eval()
    assert m.width == int(g['width']) and m.fill() == int(g['fill_stride'])
    with sim_backend.patched(), torch.no_grad():
        y = m(torch.from_numpy(g['xd'])).numpy()
    mx, l2 = rel_err(y, g['yd'])
    assert mx < TOL_SEEDED and l2 < TOL_SEEDED, (mx, l2)


def test_unet_sim_pretrained_and_seeded():
    from topaz_b200.denoising.models import UDenoiseNet, UDenoiseNet3D
    g = gold('unet_pretrained')
    m = _load(UDenoiseNet(base_width=11, top_width=5), weights_of(g)); m.eval()
    with sim_backend.patched(), torch.no_grad():
        y = m(torch.from_numpy(g['x'])).numpy()
        yo = m(torch.from_numpy(g['xo'])).numpy()
    for a, b in ((y, g['y']), (yo, g['yo'])):
        mx, l2 = rel_err(a, b)
        assert mx < 2e-3 and l2 < 2e-3, (mx, l2)
    g = gold('unet_seeded_nf16')
    m = UDenoiseNet(nf=16, base_width=7, top_width=3)
    assert list(m.state_dict().keys()) == [str(k) for k in g['keys']]
    _load(m, seeded_state(unet_shapes(16, 7, 3, 2), int(g['seed']))); m.eval()
    with sim_backend.patched(), torch.no_grad():
        y = m(torch.from_numpy(g['x'])).numpy()
    mx, l2 = rel_err(y, g['y'])
    assert mx < TOL_SEEDED and l2 < TOL_SEEDED, (mx, l2)
    g = gold('unet3d_seeded')
    m = UDenoiseNet3D(nf=48, base_width=7, top_width=3)
    assert list(m.state_dict().keys()) == [str(k) for k in g['keys']]
    _load(m, seeded_state(unet_shapes(48, 7, 3, 3), int(g['seed']))); m.eval()
    with sim_backend.patched(), torch.no_grad():
        y = m(torch.from_numpy(g['x'])).numpy()
    mx, l2 = rel_err(y, g['y'])
    assert mx < 2e-3 and l2 < 2e-3, (mx, l2)


def test_ge_binomial_step_wiring_sim():
    """GE_binomial.step through the train engine (tape, backward, flat params, Adam, optim.state) with
    simulated kernels reproduces the reference's 3-step golden."""
    import torch.nn as nn
    from topaz_b200.methods import GE_binomial
    g = gold('ge_binomial_u32'); sd = weights_of(gold('resnet8_u32_pretrained'))
    m = _load(_classifier('resnet8', 32, 1, False), sd); m.train()
    optim = torch.optim.Adam(m.parameters(), lr=2e-4)
    tr = GE_binomial(m, optim, nn.BCEWithLogitsLoss(), float(g['pi']), l2=0.0, slack=1.0)
    assert tr.header == ['loss', 'ge_penalty', 'precision', 'adjusted_precision', 'tpr', 'fpr']
    B = int(g['B'])
    Y = torch.from_numpy(g['Y'])
    outs = []
    with sim_backend.patched_training():
        for step in range(3):
            X = torch.from_numpy(np.random.default_rng(4000 + step).standard_normal((B, 71, 71)).astype(np.float32))
            if step == 0:
                from topaz_b200 import train_engine as T
                fp = T.flat_params(m)
                score = m(X).view(-1)
                dscore = torch.empty(B); out5 = torch.empty(5)
                T.ge_loss_grad(score, Y, tr.pi, tr.slack, 0, B, dscore, out5)
                T.backward(m, dscore)
                for k, p in m.named_parameters():
                    mx, l2 = rel_err(p.grad.numpy(), g['g1.' + k])
                    assert mx < 5e-4 and l2 < 5e-4, (k, mx, l2)
                fp.flat_g.zero_()
            outs.append(tr.step(X, Y))
    np.testing.assert_allclose(np.array(outs), g['outs'], rtol=5e-4, atol=1e-6)
    for k, p in m.named_parameters():
        mx, l2 = rel_err(p.detach().numpy(), g['p3.' + k])
        assert mx < 1e-4 and l2 < 1e-4, (k, mx, l2)
    st = optim.state[next(iter(m.parameters()))]
    assert float(st['step']) == 3.0 and st['exp_avg'].abs().sum() > 0
    # cached dense plans are invalidated by the in-place parameter update
    m.eval(); m.fill()
    with sim_backend.patched(), torch.no_grad():
        y = m(torch.from_numpy(gold('resnet8_u32_pretrained')['x']))
    assert torch.isfinite(y).all()


# Gradients of the first two blocks (1.5 M activations each at B = 64) are compared loosely: with He-random weights a
# handful of BatchNorm outputs lie within 1e-6 of zero, so fp32 summation-order noise flips their ReLU mask against the
# reference run, and ONE flipped element (|g| up to 8 % of max|g|: the gradient is concentrated on the 4 positive crops)
# moves those tensors by ~3e-3 rel-L2 (measured: element (61,8,10,8) of block 1).  Everything downstream of the flip
# (blocks 2-4, classifier) agrees to 2e-6, which is what pins the BatchNorm backward wiring.
BN_GRAD_TOL_EARLY = 1e-2


def _bn_training_case():
    """ResNet8(units=32, bn=True) -- the default `topaz train` model (commands/train.py:89-91) -- from seeded weights."""
    g = gold('ge_binomial_u32_bn')
    m = _classifier('resnet8', 32, 1, True)
    sd = seeded_state({k: tuple(v.shape) for k, v in m.state_dict().items()}, int(g['seed']))
    return g, _load(m, sd)


def test_ge_binomial_batchnorm_step_wiring_sim():
    """Training-mode BatchNorm (batch statistics, running-stat update, backward through bn0/bn1 of the residual blocks)
    in the train engine with simulated kernels vs the reference's 3-step golden; then eval-mode scores (running
    statistics) on crops and the filled dense forward with the folded BN."""
    import torch.nn as nn
    from topaz_b200.methods import GE_binomial
    from topaz_b200 import train_engine as T
    g, m = _bn_training_case()
    m.train()
    optim = torch.optim.Adam(m.parameters(), lr=2e-4)
    tr = GE_binomial(m, optim, nn.BCEWithLogitsLoss(), float(g['pi']), l2=0.0, slack=1.0)
    B = int(g['B']); Y = torch.from_numpy(g['Y'])
    outs = []
    with sim_backend.patched_training():
        for step in range(3):
            X = torch.from_numpy(np.random.default_rng(4000 + step).standard_normal((B, 71, 71)).astype(np.float32))
            if step == 0:       # gradient of step 1 (the BN buffers must not move twice: snapshot and restore them)
                bufs = {k: v.clone() for k, v in m.state_dict().items() if 'running' in k or 'num_batches' in k}
                fp = T.flat_params(m)
                score = m(X).view(-1)
                dscore = torch.empty(B); out5 = torch.empty(5)
                T.ge_loss_grad(score, Y, tr.pi, tr.slack, 0, B, dscore, out5)
                T.backward(m, dscore)
                for k, p in m.named_parameters():
                    mx, l2 = rel_err(p.grad.numpy(), g['g1.' + k])
                    tol = BN_GRAD_TOL_EARLY if k.startswith(('features.features.0.', 'features.features.1.')) else 1e-4
                    assert mx < 3 * tol and l2 < tol, (k, mx, l2)
                fp.flat_g.zero_()
                m.load_state_dict(bufs, strict=False)
            outs.append(tr.step(X, Y))
    np.testing.assert_allclose(np.array(outs), g['outs'], rtol=1e-3, atol=1e-6)
    for k, v in m.state_dict().items():
        if k.endswith('num_batches_tracked'):
            assert int(v) == int(g['p3.' + k]) == 3
            continue
        assert_params_after_adam(v.detach().numpy(), g['p3.' + k], 3, 5e-4, k)
    m.eval()
    with sim_backend.patched_training(), torch.no_grad():
        yc = m(torch.from_numpy(np.random.default_rng(4100).standard_normal((8, 71, 71)).astype(np.float32))).numpy()
    mx, l2 = rel_err(yc, g['y_crops'])
    assert mx < 1e-3 and l2 < 1e-3, (mx, l2)
    m.fill()
    with sim_backend.patched(), torch.no_grad():
        yd = m(torch.from_numpy(g['x_dense'])).numpy()
    mx, l2 = rel_err(yd, g['y_dense'])
    assert mx < 3e-3 and l2 < 3e-3, (mx, l2)


@pytest.mark.parametrize('tag,bn', [('ge_binomial_conv31_bn', True), ('ge_binomial_conv31_nobn', False)])
def test_ge_binomial_prelu_extractor_wiring_sim(tag, bn):
    """`topaz train -m conv31` (conv -> [BN] -> PReLU with a learnable slope, basic.py:16-78): two GE_binomial steps through
    the train engine with simulated kernels vs the reference golden -- slope gradients included."""
    import torch.nn as nn
    from topaz_b200.methods import GE_binomial
    from topaz_b200 import train_engine as T
    g = gold(tag)
    m = _classifier('conv31', 16, 2, bn)
    assert list(m.state_dict().keys()) == [str(k) for k in g['keys']]
    _load(m, seeded_state({k: tuple(v.shape) for k, v in m.state_dict().items()}, int(g['seed']))); m.train()
    assert m.width == int(g['width'])
    optim = torch.optim.Adam(m.parameters(), lr=2e-4)
    tr = GE_binomial(m, optim, nn.BCEWithLogitsLoss(), float(g['pi']))
    B = int(g['B']); Y = torch.from_numpy(g['Y'])
    outs = []
    with sim_backend.patched_training():
        for step in range(2):
            X = torch.from_numpy(np.random.default_rng(4200 + step).standard_normal((B, m.width, m.width)).astype(np.float32))
            if step == 0:
                bufs = {k: v.clone() for k, v in m.state_dict().items() if 'running' in k or 'num_batches' in k}
                fp = T.flat_params(m)
                score = m(X).view(-1)
                dscore = torch.empty(B); out5 = torch.empty(5)
                T.ge_loss_grad(score, Y, tr.pi, tr.slack, 0, B, dscore, out5)
                T.backward(m, dscore)
                errs = {k: rel_err(p.grad.numpy(), g['g1.' + k]) for k, p in m.named_parameters()}
                assert max(e[1] for e in errs.values()) < BN_GRAD_TOL_EARLY, errs
                fp.flat_g.zero_()
                m.load_state_dict(bufs, strict=False)
            outs.append(tr.step(X, Y))
    np.testing.assert_allclose(np.array(outs), g['outs'], rtol=1e-3, atol=1e-6)
    for k, v in m.state_dict().items():
        if k.endswith('num_batches_tracked'):
            assert int(v) == 2
            continue
        assert_params_after_adam(v.detach().numpy(), g['p2.' + k], 2, 5e-4, k)


def test_ge_binomial_dropout_wiring_sim():
    """`topaz train --dropout 0.25` (ResNet8, 16 units, BatchNorm): the train engine with simulated kernels, fed the keep-masks
    the reference's nn.Dropout layers drew, reproduces the reference's loss tuple and every gradient (dropout placement,
    1/(1-p) scaling, and the fused "input > 0" masks that now also cover dropped elements)."""
    import torch.nn as nn
    from common import dropout_masks_of
    from topaz_b200 import train_engine as T
    g = gold('ge_binomial_u16_dropout')
    p = float(g['p'])
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    m = LinearClassifier(get_feature_extractor('resnet8', units=16, bn=True, dropout=p))
    assert list(m.state_dict().keys()) == [str(k) for k in g['keys']]
    _load(m, seeded_state({k: tuple(v.shape) for k, v in m.state_dict().items()}, int(g['seed']))); m.train()
    B = int(g['B']); Y = torch.from_numpy(g['Y'])
    X = torch.from_numpy(np.random.default_rng(4300).standard_normal((B, 71, 71)).astype(np.float32))
    sim_backend.DROPOUT_REPLAY[:] = [torch.from_numpy(mk).permute(0, 2, 3, 1).contiguous() for mk in dropout_masks_of(g)]
    try:
        with sim_backend.patched_training():
            T.flat_params(m)
            score = m(X).view(-1)
            assert not sim_backend.DROPOUT_REPLAY            # all three masks consumed, in order
            dscore = torch.empty(B); out5 = torch.empty(5)
            T.ge_loss_grad(score, Y, float(g['pi']), 1.0, 0, B, dscore, out5)
            T.backward(m, dscore)
    finally:
        sim_backend.DROPOUT_REPLAY[:] = []
    np.testing.assert_allclose(out5.numpy(), g['out'], rtol=5e-4, atol=1e-6)
    errs = {k: rel_err(p_.grad.numpy(), g['g1.' + k]) for k, p_ in m.named_parameters()}
    assert max(e[1] for e in errs.values()) < BN_GRAD_TOL_EARLY, errs
    # eval(): dropout is the identity and the dense path builds its plan
    m.eval(); m.fill()
    with sim_backend.patched(), torch.no_grad():
        y = m(torch.zeros(1, 1, 80, 80))
    assert y.shape == (1, 1, 80, 80)


@pytest.mark.parametrize('tag', ['PN', 'PNpi', 'GE_KL', 'PU', 'PUclip'])
def test_other_objectives_wiring_sim(tag):
    """PN / GE_KL / PU drop-ins (reference methods.py:25-74,168-322) through the shared step skeleton with simulated
    kernels vs the reference's 2-step goldens."""
    import torch.nn as nn
    from topaz_b200 import methods as M
    g = gold('objectives_u32'); sd = weights_of(gold('resnet8_u32_pretrained'))
    m = _load(_classifier('resnet8', 32, 1, False), sd); m.train()
    opt = torch.optim.Adam(m.parameters(), lr=2e-4); crit = nn.BCEWithLogitsLoss()
    tr = {'PN': lambda: M.PN(m, opt, crit, pi=None), 'PNpi': lambda: M.PN(m, opt, crit, pi=0.1),
          'GE_KL': lambda: M.GE_KL(m, opt, crit, 0.035, slack=1.0, momentum=0.9),
          'PU': lambda: M.PU(m, opt, crit, 0.035, beta=0.0), 'PUclip': lambda: M.PU(m, opt, crit, 0.6, beta=0.0)}[tag]()
    B = int(g['B']); Y = torch.from_numpy(g['Y'])
    outs = []
    with sim_backend.patched_training():
        for step in range(2):
            X = torch.from_numpy(np.random.default_rng(4000 + step).standard_normal((B, 71, 71)).astype(np.float32))
            outs.append(tr.step(X, Y))
    np.testing.assert_allclose(np.array(outs), g[tag + '.outs'], rtol=5e-4, atol=1e-6)
    sdn = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    for k in ['classifier.weight', 'features.features.0.conv.weight', 'features.features.2.proj.weight', 'features.features.4.conv.bias']:
        mx, l2 = rel_err(sdn[k], g[tag + '.p.' + k])
        assert mx < 1e-4, (k, mx)
    norms = np.array([float(np.linalg.norm(v.astype(np.float64))) for v in sdn.values()])
    np.testing.assert_allclose(norms, g[tag + '.norms'], rtol=1e-5)


def test_fcnn_and_affine_denoisers_sim():
    from topaz_b200.denoising.models import DenoiseNet2, AffineDenoise
    g = gold('fcnn_affine_seeded')
    mf = DenoiseNet2(64, width=11)
    assert list(mf.state_dict().keys()) == [str(k) for k in g['keys_fcnn']]
    _load(mf, seeded_state({k: tuple(v.shape) for k, v in mf.state_dict().items()}, 301)); mf.eval()
    ma = AffineDenoise(max_size=31)
    assert list(ma.state_dict().keys()) == [str(k) for k in g['keys_affine']]
    _load(ma, seeded_state({k: tuple(v.shape) for k, v in ma.state_dict().items()}, 302)); ma.eval()
    with sim_backend.patched(), torch.no_grad():
        yf = mf(torch.from_numpy(g['x'])).numpy(); ya = ma(torch.from_numpy(g['x'])).numpy()
    mx, l2 = rel_err(yf, g['y_fcnn']); assert mx < TOL_SEEDED and l2 < TOL_SEEDED, (mx, l2)
    mx, l2 = rel_err(ya, g['y_affine']); assert mx < 1e-5, (mx, l2)


def test_preprocess_downsample_and_gmm_normalize_host_logic():
    """Operator construction of the Fourier-crop downsample, radix-select quantiles and the EM driver of the GMM
    normalisation, with the kernels replaced by the CPU simulation, against the reference goldens."""
    from topaz_b200 import preprocess, stats
    g = gold('preprocess')
    x = g['x']
    with sim_backend.patched():
        for tag, kw in [('f2', dict(factor=2)), ('f4', dict(factor=4)), ('f3', dict(factor=3)), ('s', dict(shape=(37, 50))),
                        ('f1p7', dict(factor=1.7))]:
            y = preprocess.downsample(x, **kw)
            ref = g['ds.' + tag]
            assert y.shape == ref.shape and y.dtype == np.float32
            assert np.abs(y - ref).max() < 1e-4 * np.abs(ref).max(), tag
        xt = torch.from_numpy(x)
        qs = [0.0, 0.02, 0.05, 0.1, 0.5, 0.77, 0.9, 1.0]
        np.testing.assert_allclose(stats.quantiles(xt.view(-1), qs), np.quantile(x.astype(np.float64), qs), rtol=1e-7)
        for tag in 'abc':
            img = g[f'n.{tag}.x']
            y, md = stats.normalize(img, alpha=float(g[f'n.{tag}.alpha']), beta=float(g[f'n.{tag}.beta']),
                                    num_iters=int(g[f'n.{tag}.iters']), method='gmm')
            assert abs(md['mu'] - g[f'n.{tag}.mu']) < 1e-4 * abs(md['mu']) and abs(md['std'] - g[f'n.{tag}.std']) < 1e-3 * md['std'], tag
            assert abs(md['pi'] - g[f'n.{tag}.pi']) < 1e-3
            np.testing.assert_allclose(md['logps'], g[f'n.{tag}.logps'], rtol=2e-5)
            assert y.dtype == np.float32 and np.abs(y - g[f'n.{tag}.y']).max() < 1e-3
            ya, mda = stats.normalize(img, method='affine')
            assert abs(mda['mu'] - img.astype(np.float64).mean()) < 1e-4 and mda['pi'] == 1


def test_mn_major_tf32_operand_layout_matches_the_b200_probe():
    """Hardware fact 6 as a fixture: `tests/golden/lab_mn_major_b200.json` holds, for 15 descriptor settings, the
    shared-memory word every (row, k) of an MN-major `kind::tf32` operand was READ from on a B200 (`tools/lab_mn_major.py`, one raw
    tcgen05.mma per setting against an identity operand).  The address rule the halo-resident weight-gradient kernels rely on --
    SWIZZLE_128B_BASE32B: byte = start + (m / 32) * LBO + (k / 4) * SBO + (k % 4) * 128 + (m % 32) * 4, 32-byte pieces XOR-ed with
    bits 7-8 of the address, any start row, overlapping MN atoms allowed -- reproduces every recorded address; the ordinary 128-byte
    swizzle (layout type 2) read nothing (all zeros)."""
    import json, re
    rec = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'lab_mn_major_b200.json')))
    window = 2048                                      # floats of the probe's index image that are exact in tf32
    swz32 = lambda a: a ^ (((a >> 7) & 3) << 5)
    checked = 0
    for name, r in rec.items():
        p = {k: int(v) for k, v in re.findall(r'(\w+)=(\d+)', name)}
        got = np.array(r['got'])
        if p.get('layout') == 2:
            assert (got == -1).all()                   # D was all zero: the layout is not readable MN-major for 32-bit operands
            continue
        rows = got.shape[0]                            # A: 128 rows (M), B: N rows
        hyp = np.array([[swz32(p['start'] + (m >> 5) * p['lbo'] + (k >> 2) * p['sbo'] + (k & 3) * 128 + (m & 31) * 4) // 4
                         for k in range(8)] for m in range(rows)])
        hyp[hyp >= window] = -1
        assert np.array_equal(got, hyp), name
        checked += 1
    assert checked == 15
