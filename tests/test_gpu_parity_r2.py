"""-m gpu parity tests added in round 2 (through the drop-in modules -> ctypes -> C ABI):
  * the reference's PRETRAINED networks that round 1 never loaded on the GPU: unet-3d-10a on one cfg5 patch (192^3),
    resnet16_u64 (`topaz extract`'s default), unet-small; conv127 (dilation 16 -> per-tap kernel) incl. a direct
    tpz_tc_conv_v1 call at dilation 16; conv63 u32x2 at 4096^2 (cfg2's secondary extractor) on oracle windows
  * strict mode: every He-random seeded network meets the north-star 1e-3 (the default fast mode's measured numbers are
    printed beside it)
  * fp16 range guard: 1e5-scale input, BatchNorm with running_var = 1e-12, non-representable weights raise
Fixtures: tools/make_goldens_r2.py (real reference, CPU fp32)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from common import gold, weights_of, seeded_state, rel_err, check_parity
from common_shapes import classifier_shapes, unet_shapes
from oracle import topaz_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-3
WIN3D = [(0, 0, 0), (168, 168, 168), (0, 84, 168), (84, 84, 84), (40, 120, 72)]      # tools/make_goldens_r2.py


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


class _precision:
    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        from topaz_b200 import engine
        self.old, engine.PRECISION = engine.PRECISION, self.mode

    def __exit__(self, *a):
        from topaz_b200 import engine
        engine.PRECISION = self.old


# ---------------------------------------------------------------- pretrained networks of the reference
def test_unet3d_pretrained_10a_192_patch_and_tomogram():
    """a9 / cfg5 geometry: UDenoiseNet3D (denoising/models.py:452-564) with the packaged unet-3d-10a weights on ONE 192^3
    patch of the N(0,1) tomogram, against the real reference's output (stride-8 lattice of the whole patch + five dense 24^3
    windows incl. corners), in the default precision.  Then the fast (11-bit) mode's error on the same patch is printed: it is
    the reason the default runs the last four convolutions with split operands (engine._unet_precision)."""
    from topaz_b200.denoising.models import UDenoiseNet3D
    from topaz_b200.denoise import Denoise3D
    g = gold('unet3d_pretrained_10a'); sd = weights_of(g)
    m = _load(UDenoiseNet3D(base_width=7), sd).cuda(); m.eval()
    with torch.no_grad():
        y32 = m(torch.from_numpy(g['x32']).cuda()).cpu().numpy()
    check_parity(y32, g['y32'], TOL, 'unet-3d-10a 32^3')
    x = np.random.default_rng(5000).standard_normal((192, 192, 192)).astype(np.float32)
    with torch.no_grad():
        y = m(torch.from_numpy(x).cuda()[None, None]).cpu().numpy()[0, 0]
    assert np.isfinite(y).all()
    scale = float(g['y192_stats'][2])                      # max|ref| over the whole patch
    lat = g['y192_lattice']
    d = np.abs(y[::8, ::8, ::8] - lat)
    print(f'unet-3d-10a 192^3 lattice: max-rel {d.max() / scale:.2e} rel-L2 {np.linalg.norm(d) / np.linalg.norm(lat):.2e}')
    assert d.max() / scale < TOL and np.linalg.norm(d) / np.linalg.norm(lat) < TOL
    for i, (a, b, c) in enumerate(WIN3D):
        ref = g[f'y192_win{i}']
        got = y[a:a + 24, b:b + 24, c:c + 24]
        mx, l2 = float(np.abs(got - ref).max() / scale), float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        print(f'unet-3d-10a 192^3 window {i} at {(a, b, c)}: max-rel {mx:.2e} rel-L2 {l2:.2e}')
        assert mx < TOL and l2 < TOL, (i, mx, l2)
    with _precision('fast'), torch.no_grad():
        yf = m(torch.from_numpy(x).cuda()[None, None]).cpu().numpy()[0, 0]
    df = np.abs(yf[::8, ::8, ::8] - lat)
    print(f'unet-3d-10a 192^3 in fast mode (11-bit operands, = TF32 of the reference GPU path): max-rel {df.max() / scale:.2e} '
          f'rel-L2 {np.linalg.norm(df) / np.linalg.norm(lat):.2e}')
    # Denoise3D.denoise (denoise.py:336-377) with the pretrained model, DC-free metric (tomo = 5 + 2 N(0,1))
    d3 = Denoise3D(m)
    yt = d3.denoise(g['tomo'].copy(), patch_size=32, padding=16, verbose=False)
    check_parity(yt, g['y_tomo'], TOL, 'Denoise3D pretrained 70x50x64', dc_free=True)


def test_resnet16_u64_pretrained_dense():
    """`topaz extract`'s default model (commands/extract.py:18 -> factory.py:34-36) with its packaged weights."""
    g = gold('resnet16_u64_pretrained'); sd = weights_of(g)
    m = _load(_classifier('resnet16', 64), sd).cuda(); m.eval()
    assert m.fill() == int(g['fill_stride'])
    with torch.no_grad():
        y = m(torch.from_numpy(g['x']).cuda()).cpu().numpy()
    check_parity(y, g['y_dense'], TOL, 'resnet16_u64 dense')
    x = np.random.default_rng(16).standard_normal((1, 1, 300, 257)).astype(np.float32)
    with torch.no_grad():
        y = m(torch.from_numpy(x).cuda()).cpu().numpy()
    check_parity(y, O.classifier_forward(sd, x, 'resnet16', 64, filled=True).numpy(), TOL, 'resnet16_u64 300x257 vs oracle')


def test_unet_small_pretrained():
    from topaz_b200.denoising.models import UDenoiseNetSmall
    from topaz_b200.denoise import Denoise
    g = gold('unet_small_pretrained'); sd = weights_of(g)
    m = _load(UDenoiseNetSmall(width=11, top_width=5), sd).cuda(); m.eval()
    with torch.no_grad():
        y = m(torch.from_numpy(g['x']).cuda()).cpu().numpy()
        yo = m(torch.from_numpy(g['xo']).cuda()).cpu().numpy()
    check_parity(y, g['y'], TOL, 'unet-small'); check_parity(yo, g['yo'], TOL, 'unet-small odd sizes')
    dn = Denoise(m)
    check_parity(dn._denoise(g['img'].copy()), g['y_call'], TOL, 'unet-small _denoise', dc_free=True)
    check_parity(dn.denoise(g['img'].copy(), patch_size=64, padding=24), g['y_pat'], TOL, 'unet-small patched', dc_free=True)


def test_conv127_dilation16_dense_and_direct_v1_kernel():
    """conv127 filled reaches dilation 16 in its last layer: outside the halo-resident kernel's lattice range (<= 8), so the
    dispatcher takes the per-tap kernel (tpz_tc_conv_v1).  Whole network vs the reference golden, then ONE dilation-16 layer
    through tpz_tc_conv_v1 directly against an fp32 torch convolution of the same fp16-rounded operands."""
    from topaz_b200 import ops
    g = gold('cls_conv127_u16x2')
    m = _load(_classifier('conv127', 16, 2, True), seeded_state(classifier_shapes('conv127', 16, 2, True), int(g['seed'])))
    m.cuda(); m.eval()
    assert m.fill() == int(g['fill_stride'])
    with torch.no_grad():
        y = m(torch.from_numpy(g['xd']).cuda()).cpu().numpy()
    mx, l2 = check_parity(y, g['yd'], 3e-3, 'conv127 dense (fast mode, seeded weights)')
    with _precision('strict'), torch.no_grad():
        ys = m(torch.from_numpy(g['xd']).cuda()).cpu().numpy()
    check_parity(ys, g['yd'], TOL, 'conv127 dense (strict mode)')
    gen = torch.Generator().manual_seed(16)
    for (ci, co, H, W) in [(64, 128, 90, 140), (32, 64, 70, 66)]:
        w = torch.randn(co, ci, 5, 5, generator=gen) * (2.0 / (ci * 25)) ** 0.5
        b = torch.randn(co, generator=gen) * 0.1
        a = torch.randn(1, 1, H, W, ci, generator=gen).half()
        plan = ops.pack_tc_conv([ops.ConvPart(w, ci, 16)], b, co, 0.25, 'cuda')
        Ho, Wo = H - 64, W - 64
        out = torch.empty((1, 1, Ho, Wo, co), dtype=torch.float16, device='cuda')
        old = ops.TC_VARIANT
        try:
            ops.TC_VARIANT = 'v1'
            ops.tc_conv(plan, [a.cuda()], (1, 1, Ho, Wo), out=out)
        finally:
            ops.TC_VARIANT = old
        ref = F.conv2d(a[0, 0].permute(2, 0, 1)[None].float(), w.half().float(), b, dilation=16)[0]
        ref = torch.where(ref > 0, ref, ref * 0.25).permute(1, 2, 0)
        check_parity(out[0, 0].float().cpu().numpy(), ref.numpy(), TOL, f'tpz_tc_conv_v1 dilation 16 {ci}->{co}')


def test_conv63_u32x2_4096_windows_vs_oracle():
    """cfg2's secondary extractor: conv63 (units 32, scaling 2, BN, PReLU; basic.py:12-111) over one 4096^2 micrograph, checked
    against the oracle on windows cut with their receptive-field ring (width 63 -> 31 px), as for ResNet8 in test_gpu_fullsize.
    Seeded He-random weights (no packaged conv63): fast-mode tolerance 3e-3, strict mode 1e-3 on the same windows."""
    from topaz_b200.extract import score_arrays
    g = gold('cls_conv63_u32x2')
    sd = seeded_state(classifier_shapes('conv63', 32, 2, True), int(g['seed']))
    S, halo, win = 4096, 31, 96
    x = np.random.default_rng(1001).standard_normal((S, S)).astype(np.float32)
    xp = np.pad(x, halo)
    corners = [(0, 0), (S - win, S - win), (1777, 3000)]
    refs = [O.classifier_forward(sd, xp[i:i + win + 2 * halo, j:j + win + 2 * halo][None, None], 'conv63', 32, True, True, 2)
            .numpy()[0, 0][halo:halo + win, halo:halo + win] for (i, j) in corners]
    for mode, tol in (('auto', 3e-3), ('strict', TOL)):
        m = _load(_classifier('conv63', 32, 2, True), sd); m.eval()
        with _precision(mode):
            (y,) = list(score_arrays(m, [x]))
        assert y.shape == (S, S) and np.isfinite(y).all()
        scale = np.abs(y).max()
        for (i, j), ref in zip(corners, refs):
            got = y[i:i + win, j:j + win]
            mx, l2 = float(np.abs(got - ref).max() / scale), float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
            print(f'conv63 4096^2 [{mode}] window {(i, j)}: max-rel {mx:.2e} rel-L2 {l2:.2e}')
            assert mx < tol and l2 < tol, (mode, i, j, mx, l2)


# ---------------------------------------------------------------- strict mode
@pytest.mark.parametrize('name,arch,units,scaling,bn', [
    ('resnet16_u16', 'resnet16', 16, 1, False),
    ('resnet8_u16_bn', 'resnet8', 16, 1, True),
    ('conv31_u16x2', 'conv31', 16, 2, True),
    ('conv63_u32x2', 'conv63', 32, 2, True),
    ('conv63_u16_nobn', 'conv63', 16, 1, False),
])
def test_strict_mode_seeded_classifiers_meet_1e3(name, arch, units, scaling, bn):
    g = gold('cls_' + name)
    m = _load(_classifier(arch, units, scaling, bn), seeded_state(classifier_shapes(arch, units, scaling, bn), int(g['seed'])))
    m.cuda(); m.eval(); m.fill()
    with torch.no_grad():
        yf = m(torch.from_numpy(g['xd']).cuda()).cpu().numpy()
    with _precision('strict'), torch.no_grad():
        ys = m(torch.from_numpy(g['xd']).cuda()).cpu().numpy()
    mxf, l2f = rel_err(yf, g['yd'])
    print(f'{name}: fast mode max-rel {mxf:.2e} rel-L2 {l2f:.2e}')
    check_parity(ys, g['yd'], TOL, f'{name} strict mode')


def test_strict_mode_seeded_unets_and_fcnn_meet_1e3():
    from topaz_b200.denoising.models import UDenoiseNet, UDenoiseNet3D, DenoiseNet2
    cases = []
    g = gold('unet_seeded_nf16')
    cases.append(('unet nf16', _load(UDenoiseNet(nf=16, base_width=7, top_width=3), seeded_state(unet_shapes(16, 7, 3, 2), int(g['seed']))), g['x'], g['y']))
    g = gold('unet3d_seeded')
    cases.append(('unet3d seeded', _load(UDenoiseNet3D(nf=48, base_width=7, top_width=3), seeded_state(unet_shapes(48, 7, 3, 3), int(g['seed']))), g['x'], g['y']))
    g = gold('fcnn_affine_seeded')
    mf = DenoiseNet2(64, width=11)
    cases.append(('fcnn', _load(mf, seeded_state({k: tuple(v.shape) for k, v in mf.state_dict().items()}, 301)), g['x'], g['y_fcnn']))
    g = gold('unet_pretrained')
    cases.append(('unet pretrained odd sizes', _load(UDenoiseNet(base_width=11, top_width=5), weights_of(g)), g['xo'], g['yo']))
    for name, m, x, ref in cases:
        m.cuda(); m.eval()
        with _precision('fast'), torch.no_grad():
            yf = m(torch.from_numpy(x).cuda()).cpu().numpy()
        with _precision('strict'), torch.no_grad():
            ys = m(torch.from_numpy(x).cuda()).cpu().numpy()
        mxf, l2f = rel_err(yf, ref)
        print(f'{name}: fast mode max-rel {mxf:.2e} rel-L2 {l2f:.2e}')
        check_parity(ys, ref, TOL, f'{name} strict mode')


# ---------------------------------------------------------------- fp16 range guard
def test_range_guard_unnormalised_input_and_bn_tiny_running_var():
    """An un-normalised micrograph (1e5 * N(0,1): |x| far beyond the fp16 maximum 65504) is scored like the fp32 reference:
    the activations are stored multiplied by a power of two chosen on the device from max|x| (tpz_range_scale), the biases are
    scaled with them, the fused classifier output is scaled back.  Same for the U-Net called directly on raw-scale input, and
    for a BatchNorm layer with running_var = 1e-12 (folded rows 316x larger)."""
    from topaz_b200 import ops
    from topaz_b200.denoising.models import UDenoiseNet
    g = gold('resnet8_u32_pretrained'); sd = weights_of(g)
    m = _load(_classifier('resnet8', 32), sd).cuda(); m.eval(); m.fill()
    for scale in (1e5, 3e-6, 1.0, 7e8):
        x = (g['x'] * np.float32(scale)).astype(np.float32)
        with torch.no_grad():
            y = m(torch.from_numpy(x).cuda()).cpu().numpy()
        assert np.isfinite(y).all(), scale
        check_parity(y, O.classifier_forward(sd, x, 'resnet8', 32, filled=True).numpy(), TOL, f'range guard x{scale:g}')
    rng = ops.range_scale(torch.from_numpy(g['x'] * np.float32(1e5)).cuda()).cpu().numpy()
    assert rng[0] < 1 and rng[0] * rng[1] == 1.0 and np.log2(rng[0]) == np.round(np.log2(rng[0]))
    assert 4 <= np.abs(g['x']).max() * 1e5 * rng[0] < 8
    assert ops.range_scale(torch.from_numpy(g['x']).cuda()).cpu().tolist() == [1.0, 1.0]
    gu = gold('unet_pretrained'); sdu = weights_of(gu)
    mu = _load(UDenoiseNet(base_width=11, top_width=5), sdu).cuda(); mu.eval()
    x = (gu['x'] * np.float32(2e5)).astype(np.float32)
    with torch.no_grad():
        y = mu(torch.from_numpy(x).cuda()).cpu().numpy()
    check_parity(y, O.unet_forward(sdu, x).numpy(), TOL, 'range guard unet x2e5')
    gb = gold('cls_resnet8_u16_bn')
    sdb = seeded_state(classifier_shapes('resnet8', 16, 1, True), int(gb['seed']))
    sdb['features.features.1.bn1.running_var'][:] = 1e-12
    mb = _load(_classifier('resnet8', 16, 1, True), sdb).cuda(); mb.eval(); mb.fill()
    with torch.no_grad():
        y = mb(torch.from_numpy(gb['xd']).cuda()).cpu().numpy()
    assert np.isfinite(y).all()
    check_parity(y, O.classifier_forward(sdb, gb['xd'], 'resnet8', 16, filled=True, bn=True).numpy(), 3e-3, 'BN running_var=1e-12')
    sdb['features.features.1.bn1.weight'][:] = np.inf
    mb = _load(_classifier('resnet8', 16, 1, True), sdb).cuda(); mb.eval(); mb.fill()
    with pytest.raises(RuntimeError), torch.no_grad():
        mb(torch.from_numpy(gb['xd']).cuda())


def test_split_maxpool_and_first_layer_kernels():
    """strict-mode building blocks vs torch: tpz_maxpool2 over (hi, lo) pairs, tpz_conv_first / tpz_im2col(3d)_first with split
    outputs (hi + lo reproduces the fp32 value to 2^-22)."""
    from topaz_b200 import ops
    gen = torch.Generator().manual_seed(3)
    v = torch.randn(1, 4, 10, 14, 32, generator=gen) * 3
    hi = v.half(); lo = (v - hi.float()).half()
    x = torch.cat([hi, lo], -1).cuda()
    for dims in (2, 3):
        y = ops.maxpool2(x, dims, split=True).float().cpu()
        ref = F.max_pool3d((hi.float() + lo.float()).permute(0, 4, 1, 2, 3), (2 if dims == 3 else 1, 2, 2)).permute(0, 2, 3, 4, 1)
        got = y[..., :32] + y[..., 32:]
        assert float((got - ref).abs().max()) <= 2.0 ** -20 * float(ref.abs().max())
    img = torch.randn(2, 40, 52, generator=gen)
    w = torch.randn(48, 1, 7, 7, generator=gen) / 7
    b = torch.randn(48, generator=gen)
    y = ops.conv_first(img[:, None].cuda(), w.cuda(), b.cuda(), 1, 3, 0.1, 64, split=True).float().cpu()
    ref = F.leaky_relu(F.conv2d(img[:, None], w, b, padding=3), 0.1).permute(0, 2, 3, 1)
    got = y[:, 0, :, :, :48] + y[:, 0, :, :, 64:112]
    assert float((got - ref).abs().max()) < 1e-5 * float(ref.abs().max())
    col = ops.im2col_first(img.cuda(), 5, 2, 32, split=True).float().cpu()
    refc = F.unfold(img[:, None], 5, padding=2).reshape(2, 25, 40, 52).permute(0, 2, 3, 1)
    got = col[:, 0, :, :, :25] + col[:, 0, :, :, 32:57]
    assert float((got - refc).abs().max()) <= 2.0 ** -20 * float(refc.abs().max())
    vol = torch.randn(1, 9, 12, 10, generator=gen)
    col3 = ops.im2col3d_first(vol.cuda(), 3, 32, split=True).float().cpu()
    ref3 = F.pad(vol, (1,) * 6)
    t = 0
    for dz in range(3):
        for dy in range(3):
            for dx in range(3):
                r = ref3[:, dz:dz + 9, dy:dy + 12, dx:dx + 10]
                assert float((col3[..., t] + col3[..., 32 + t] - r).abs().max()) <= 2.0 ** -20 * 5
                t += 1


def test_conv_last3d_tiled_kernel_plain_and_split_inputs():
    """tpz_conv_last 3x3x3 (z-marching CUDA-core kernel) vs torch conv3d: plain fp16 input (C = 32) and split (hi, lo) input
    as 64 channels with the weights repeated; ragged sizes, batch 2, fused de-normalisation."""
    from topaz_b200 import ops
    gen = torch.Generator().manual_seed(9)
    w = torch.randn(27, 32, generator=gen) * 0.1
    stats = torch.tensor([0.7, 1.9])
    for (N, D, H, W) in [(1, 20, 37, 45), (2, 5, 16, 32), (1, 33, 70, 9)]:
        v = torch.randn(N, D, H, W, 32, generator=gen) * 2
        hi = v.half(); lo = (v - hi.float()).half()
        ref_w = w.t().reshape(1, 32, 3, 3, 3)
        ref_hi = F.conv3d(hi.float().permute(0, 4, 1, 2, 3), ref_w, None, padding=1)[:, 0] + 0.3
        ref_full = F.conv3d((hi.float() + lo.float()).permute(0, 4, 1, 2, 3), ref_w, None, padding=1)[:, 0] + 0.3
        y = ops.conv_last(hi.cuda(), 32, w.cuda(), 0.3, (3, 3, 3), 1, 1).cpu()
        check_parity(y.numpy(), ref_hi.numpy(), 1e-5, f'conv_last3d plain {N}x{D}x{H}x{W}')
        x2 = torch.cat([hi, lo], -1).cuda()
        y2 = ops.conv_last(x2, 32, torch.cat([w, w], 1).cuda(), 0.3, (3, 3, 3), 1, 1, stats=stats.cuda()).cpu()
        check_parity(y2.numpy(), (ref_full * stats[1] + stats[0]).numpy(), 1e-5, f'conv_last3d split {N}x{D}x{H}x{W}')


def test_training_step_graph_replay_matches_eager():
    """GE_binomial.step replayed from a CUDA graph (third step onwards) gives the same parameters as the eager path
    (TPZ_TRAIN_GRAPH=0) after 6 steps, incl. the device-side Adam step count."""
    import os
    import torch.nn as nn
    from topaz_b200.methods import GE_binomial
    sd = weights_of(gold('resnet8_u32_pretrained'))
    Y = torch.tensor([1.0] * 4 + [0.0] * 60, dtype=torch.float64).cuda()
    Xs = [torch.from_numpy(np.random.default_rng(4100 + s).standard_normal((64, 71, 71)).astype(np.float32)).cuda() for s in range(6)]
    res = {}
    old = os.environ.get('TPZ_TRAIN_GRAPH')
    try:
        for mode in ('0', '1'):
            os.environ['TPZ_TRAIN_GRAPH'] = mode
            m = _load(_classifier('resnet8', 32), sd).cuda(); m.train()
            tr = GE_binomial(m, torch.optim.Adam(m.parameters(), lr=2e-4), nn.BCEWithLogitsLoss(), 0.035)
            outs = [tr.step(X, Y) for X in Xs]
            st = tr.__dict__.get('_graph_state', {})
            assert (st.get('graph') is not None) == (mode == '1'), (mode, st.get('key'))
            res[mode] = (np.array(outs), torch.cat([p.detach().reshape(-1) for p in m.parameters()]).cpu().numpy(),
                         tr.optim.state[next(iter(m.parameters()))]['step'].item())
    finally:
        if old is None:
            os.environ.pop('TPZ_TRAIN_GRAPH', None)
        else:
            os.environ['TPZ_TRAIN_GRAPH'] = old
    assert res['0'][2] == res['1'][2] == 6.0
    np.testing.assert_allclose(res['1'][0], res['0'][0], rtol=2e-5, atol=1e-7)
    mx, l2 = rel_err(res['1'][1], res['0'][1])
    assert mx < 1e-5 and l2 < 1e-5, (mx, l2)          # wgrad atomics reorder fp32 sums from run to run
