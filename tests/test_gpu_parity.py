"""-m gpu parity tests: the CUDA path (through the drop-in modules -> ctypes -> C ABI) against the
reference goldens and the CPU oracle.  Tolerance: north-star 1e-3 (max|d|/max|ref| and rel-L2) for every
network with the reference's PRETRAINED weights in the default precision; He-random seeded weights amplify the
2^-11 operand rounding of the default (fast) mode through 9-16 layers to 1-3e-3 -- those cases are held to 3e-3
here and to 1e-3 in strict mode by tests/test_gpu_parity_r2.py::test_strict_mode_*.  De-normalised denoiser
outputs (mean 10, std 0.1) use the DC-free form of the metric (common.rel_err_dc).  Every check prints the
metric and the worst element-wise relative error on |ref| > 0.1 (run pytest with -s / -rP to see them)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from common import gold, weights_of, seeded_state, rel_err, check_parity
from common_shapes import classifier_shapes, unet_shapes
from oracle import topaz_oracle as O

pytestmark = pytest.mark.gpu
TOL, TOL_SEEDED = 1e-3, 2e-3      # measured on B200 in the default mode: seeded nets 4.9e-4 ... 1.7e-3 (GPUTEST log, profiles/)


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


def _check(y, ref, tol, what='', dc_free=False):
    import inspect
    what = what or inspect.stack()[1].function
    return check_parity(y, ref, tol, what, dc_free=dc_free)


def test_native_library_loaded_and_device_is_sm100():
    import ctypes as C
    from topaz_b200 import _lib
    sms, ma, mi = C.c_int(), C.c_int(), C.c_int()
    _lib.check(_lib.lib().tpz_device_info(C.byref(sms), C.byref(ma), C.byref(mi)))
    assert ma.value == 10 and sms.value >= 100, (sms.value, ma.value, mi.value)


# ---------------------------------------------------------------- direct kernels
def test_conv_first_matches_torch():
    from topaz_b200 import ops
    g = torch.Generator().manual_seed(1)
    for (N, D, H, W, co, kd, k, pad, slope) in [(2, 1, 70, 90, 64, 1, 7, 35, 0.0), (1, 1, 65, 130, 48, 1, 11, 5, 0.1),
                                                (1, 20, 24, 40, 48, 7, 7, 3, 0.1)]:
        x = torch.randn(N, D, H, W, generator=g)
        w = torch.randn(co, kd, k, k, generator=g) * 0.1
        b = torch.randn(co, generator=g) * 0.1
        out = ops.conv_first(x.cuda(), w.cuda(), b.cuda(), 1, pad, slope, 64).cpu().float()
        if kd > 1:
            ref = F.conv3d(x[:, None], w[:, None], b, padding=pad)
        else:
            ref = F.conv2d(x.reshape(N, 1, H, W), w, b, padding=pad)[:, :, None]
        ref = torch.where(ref > 0, ref, ref * slope).permute(0, 2, 3, 4, 1)
        _check(out[..., :co], ref, 1e-3)
        assert (out[..., co:] == 0).all()


def test_pool_upsample_last_meanstd():
    from topaz_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, 5, 21, 27, 64, generator=g).half()
    for dims in (2, 3):
        y = ops.maxpool2(x.cuda(), dims).cpu().float()
        ref = F.max_pool3d(x.float().permute(0, 4, 1, 2, 3), (2 if dims == 3 else 1, 2, 2)).permute(0, 2, 3, 4, 1)
        assert torch.equal(y, ref)
    xs = torch.randn(1, 3, 10, 13, 32, generator=g).half()
    for size in [(3, 21, 27), (7, 20, 26), (6, 95, 77)]:
        y = ops.upsample_nearest(xs.cuda(), size).cpu().float()
        ref = F.interpolate(xs.float().permute(0, 4, 1, 2, 3), size=size, mode='nearest').permute(0, 2, 3, 4, 1)
        assert torch.equal(y, ref), size
    w = torch.randn(27, 32, generator=g) * 0.1
    y = ops.conv_last(xs.cuda(), 32, w.cuda(), 0.3, (3, 3, 3), 1, 1).cpu()
    ref = F.conv3d(xs.float().permute(0, 4, 1, 2, 3), w.t().reshape(1, 32, 3, 3, 3), None, padding=1)[:, 0] + 0.3
    _check(y, ref, 1e-4)
    v = (10 + 3 * torch.randn(1000003, generator=g))
    for unb in (True, False):
        st = ops.meanstd(v.cuda(), unb).cpu()
        assert abs(st[0] - v.double().mean()) < 1e-5 and abs(st[1] - v.double().std(unbiased=unb)) < 1e-5
    st = ops.meanstd(v.cuda(), True)
    n = ops.affine(v.cuda(), st).cpu()
    _check(n, (v - v.mean()) / v.std(), 1e-5)


# ---------------------------------------------------------------- classifiers
def test_resnet8_u32_pretrained_dense():
    g = gold('resnet8_u32_pretrained'); sd = weights_of(g)
    m = _load(_classifier('resnet8', 32), sd).cuda(); m.eval(); m.fill()
    with torch.no_grad():
        y = m(torch.from_numpy(g['x']).cuda()).cpu().numpy()
    _check(y, g['y_dense'], TOL)
    _check(y, O.classifier_forward(sd, g['x'], 'resnet8', 32, filled=True).numpy(), TOL)


def test_resnet8_u64_pretrained_dense_and_larger_image_vs_oracle():
    g = gold('resnet8_u64_pretrained'); sd = weights_of(g)
    m = _load(_classifier('resnet8', 64), sd).cuda(); m.eval(); m.fill()
    with torch.no_grad():
        y = m(torch.from_numpy(g['x']).cuda()).cpu().numpy()
    _check(y, g['y_dense'], TOL)
    x = np.random.default_rng(1234).standard_normal((2, 1, 200, 333)).astype(np.float32)   # ragged sizes, batch 2
    with torch.no_grad():
        y = m(torch.from_numpy(x).cuda()).cpu().numpy()
    _check(y, O.classifier_forward(sd, x, 'resnet8', 64, filled=True).numpy(), TOL)


@pytest.mark.parametrize('name,arch,units,scaling,bn', [
    ('resnet16_u16', 'resnet16', 16, 1, False),
    ('resnet8_u16_bn', 'resnet8', 16, 1, True),
    ('conv31_u16x2', 'conv31', 16, 2, True),
    ('conv63_u32x2', 'conv63', 32, 2, True),
    ('conv63_u16_nobn', 'conv63', 16, 1, False),
])
def test_seeded_classifiers_dense(name, arch, units, scaling, bn):
    g = gold('cls_' + name)
    m = _load(_classifier(arch, units, scaling, bn), seeded_state(classifier_shapes(arch, units, scaling, bn), int(g['seed'])))
    m.cuda(); m.eval()
    assert m.fill() == int(g['fill_stride'])
    with torch.no_grad():
        y = m(torch.from_numpy(g['xd']).cuda()).cpu().numpy()
    _check(y, g['yd'], TOL_SEEDED)


def test_score_images_and_patched_scoring(tmp_path):
    from topaz_b200 import mrc
    from topaz_b200.extract import score_images
    g = gold('resnet8_u32_pretrained'); sd = weights_of(g)
    m = _load(_classifier('resnet8', 32), sd)
    paths = []
    for i, arr in enumerate([g['xp'][0, 0], g['x'][0, 0]]):
        p = str(tmp_path / f'mic{i}.mrc'); mrc.write(p, arr); paths.append(p)
    outs = list(score_images(m, paths, device=0))
    assert [p for p, _ in outs] == paths and outs[0][1].dtype == np.float32
    _check(outs[0][1], g['y_full'][0, 0], TOL)
    _check(outs[1][1], g['y_dense'][0, 0], TOL)
    outs = list(score_images(m, paths[:1], device=0, patch_size=64))
    assert outs[0][1].dtype == np.float64
    _check(outs[0][1], g['y_patch'][0, 0], TOL)
    # bit-exact coordinate extraction when both NMS implementations see the same score map is a host-side
    # property (tests/test_oracle_golden.py); here: the GPU score map picks the same top particles
    s_ref, c_ref = O.nms(g['y_full'][0, 0], 6, -6.0)
    s_gpu, c_gpu = O.nms(list(score_images(m, paths[:1], device=0))[0][1], 6, -6.0)
    k = min(20, len(c_ref))
    top_gpu = {tuple(c) for c in c_gpu[:k + 5]}
    assert all(tuple(c) in top_gpu for c in c_ref[:k])


# ---------------------------------------------------------------- denoisers
def test_unet_pretrained_forward_and_pipeline():
    from topaz_b200.denoising.models import UDenoiseNet
    from topaz_b200.denoise import Denoise
    g = gold('unet_pretrained'); sd = weights_of(g)
    m = _load(UDenoiseNet(base_width=11, top_width=5), sd).cuda(); m.eval()
    with torch.no_grad():
        y = m(torch.from_numpy(g['x']).cuda()).cpu().numpy()
        yo = m(torch.from_numpy(g['xo']).cuda()).cpu().numpy()
    _check(y, g['y'], TOL); _check(yo, g['yo'], TOL)
    dn = Denoise(m)
    # de-normalised outputs are mean 10 / std 0.1: DC-free metric (relative to the denoised SIGNAL, not to its offset)
    _check(dn._denoise(g['img'].copy()), g['y_call'], TOL, dc_free=True)
    _check(dn.denoise(g['img'].copy(), patch_size=64, padding=24), g['y_pat'], TOL, dc_free=True)


def test_unet_seeded_2d_and_3d():
    from topaz_b200.denoising.models import UDenoiseNet, UDenoiseNet3D
    from topaz_b200.denoise import Denoise3D
    g = gold('unet_seeded_nf16')
    m = _load(UDenoiseNet(nf=16, base_width=7, top_width=3), seeded_state(unet_shapes(16, 7, 3, 2), int(g['seed']))).cuda()
    with torch.no_grad():
        y = m(torch.from_numpy(g['x']).cuda()).cpu().numpy()
    _check(y, g['y'], TOL_SEEDED)
    g = gold('unet3d_seeded')
    m = _load(UDenoiseNet3D(nf=48, base_width=7, top_width=3), seeded_state(unet_shapes(48, 7, 3, 3), int(g['seed']))).cuda()
    with torch.no_grad():
        y = m(torch.from_numpy(g['x']).cuda()).cpu().numpy()
    _check(y, g['y'], TOL_SEEDED)
    d3 = Denoise3D(m)
    yt = d3.denoise(g['tomo'].copy(), patch_size=16, padding=8, verbose=False)
    _check(yt, g['y_tomo'], TOL_SEEDED, dc_free=True)
    # patch sharding (multi-GPU partition of the patch list) reassembles to the same volume
    n = int(np.prod([int(np.ceil(s / 16)) for s in g['tomo'].shape]))
    parts = [d3.denoise(g['tomo'].copy(), 16, 8, verbose=False, patch_range=(a, b)) for a, b in ((0, n // 2), (n // 2, n))]
    assert np.array_equal(parts[0] + parts[1], yt)


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()


@pytest.mark.parametrize('shape', [(1, 1, 5, 7), (3, 1, 33, 9), (1, 1, 1, 64), (2, 1, 130, 257)])
def test_dense_classifier_ragged_and_tiny_images(shape):
    """Edge geometry: images smaller than one 8x32 lattice tile / the receptive field, 1-pixel strips, batches > 1."""
    g = gold('resnet8_u32_pretrained'); sd = weights_of(g)
    m = _load(_classifier('resnet8', 32), sd).cuda(); m.eval(); m.fill()
    x = np.random.default_rng(7).standard_normal(shape).astype(np.float32)
    with torch.no_grad():
        y = m(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = O.classifier_forward(sd, x, 'resnet8', 32, filled=True).numpy()
    assert y.shape == ref.shape == shape
    _check(y, ref, TOL)


@pytest.mark.parametrize('shape', [(1, 1, 32, 32), (2, 1, 40, 72), (1, 1, 33, 47), (1, 1, 250, 130),
                                   (1, 1, 96, 40)])      # 5 x 3 = 15 tiles: odd count -> the CTA pair's repeat-tile path
def test_unet_edge_sizes(shape):
    """Smallest legal U-Net inputs (5 poolings), odd sizes (non-2x up-sampling -> gather path) and even sizes (fused path)."""
    from topaz_b200.denoising.models import UDenoiseNet
    g = gold('unet_pretrained'); sd = weights_of(g)
    m = _load(UDenoiseNet(base_width=11, top_width=5), sd).cuda(); m.eval()
    x = np.random.default_rng(11).standard_normal(shape).astype(np.float32)
    with torch.no_grad():
        y = m(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = O.unet_forward(sd, x).numpy()
    _check(y, ref, TOL)


def test_variants_agree_v1_v2():
    """The per-tap kernel (v1) and the halo-resident kernel (v2) compute the same layer outputs (fp32 accumulation order
    differs only inside the tensor core)."""
    from topaz_b200 import ops
    g = gold('resnet8_u64_pretrained'); sd = weights_of(g)
    m = _load(_classifier('resnet8', 64), sd).cuda(); m.eval(); m.fill()
    x = torch.from_numpy(np.random.default_rng(3).standard_normal((1, 1, 160, 200)).astype(np.float32)).cuda()
    outs = {}
    try:
        for v in ('v2', 'v1'):
            ops.TC_VARIANT = v
            with torch.no_grad():
                outs[v] = m(x).cpu().numpy()
    finally:
        ops.TC_VARIANT = 'auto'
    _check(outs['v1'], outs['v2'], 1e-4)


def test_filters_and_affine_normalize_dropins():
    """GaussianDenoise.apply (filters.py:62-79) and stats.normalize(method='affine') (stats.py:36-46) vs reference goldens."""
    from topaz_b200.filters import GaussianDenoise
    from topaz_b200.stats import normalize
    g = gold('filters')
    _check(GaussianDenoise(1.5).apply(g['img'].copy()), g['gauss'], 1e-5)
    n, meta = normalize(g['img'].copy(), method='affine')
    _check(n, g['norm'], 1e-5)
    assert abs(meta['mu'] - float(g['mu'])) < 1e-4 and abs(meta['std'] - float(g['std'])) < 1e-4 and n.dtype == np.float32
    vol = np.random.default_rng(5).standard_normal((12, 20, 16)).astype(np.float32)
    _check(GaussianDenoise(0.8, dims=3).apply(vol), O.gaussian_denoise(vol, 0.8), 1e-5)


def test_denoise_image_pipeline():
    """denoise_image (denoise.py:382-416): numpy normalise -> patched denoise -> restore scale, vs the oracle composition."""
    from topaz_b200.denoising.models import UDenoiseNet
    from topaz_b200.denoise import Denoise, denoise_image
    g = gold('unet_pretrained'); sd = weights_of(g)
    dn = Denoise(_load(UDenoiseNet(base_width=11, top_width=5), sd))
    mic = g['img']
    mu, std = mic.mean(), mic.std()
    ref = std * O.denoise(sd, (mic - mu) / std, patch_size=64, padding=24) + mu
    _check(denoise_image(mic.copy(), [dn], patch_size=64, padding=24), ref, TOL, dc_free=True)
    refn = O.denoise(sd, (mic - mu) / std, patch_size=64, padding=24)
    refn = (refn - refn.mean()) / refn.std()
    _check(denoise_image(mic.copy(), [dn, dn], patch_size=64, padding=24, normalize=True), refn, TOL)


@pytest.mark.parametrize('H,W,r,thr', [(20, 24, 3, -np.inf), (25, 25, 2, -np.inf), (40, 6, 3, -np.inf), (6, 40, 3, -np.inf),
                                       (33, 35, 8, -2.0), (300, 420, 9, -1.0), (8, 8, 1, 0.5)])
def test_gpu_nms_bit_exact(H, W, r, thr):
    """GPU greedy NMS vs the oracle restatement of topaz.algorithms.non_maximum_suppression (pinned to the reference
    by tests/test_oracle_golden.py): identical scores and (x, y) coordinates, incl. the right-border clip quirk."""
    from topaz_b200.algorithms import non_maximum_suppression
    x = np.random.default_rng(H * 1000 + W).standard_normal((H, W)).astype(np.float32)
    s_ref, c_ref = O.nms(x, r, thr)
    s, c = non_maximum_suppression(x, r, thr)
    assert s.dtype == np.float32 and c.dtype == np.int32
    assert np.array_equal(c, c_ref) and np.array_equal(s, s_ref)


def test_gpu_nms_on_reference_score_map():
    from topaz_b200.algorithms import non_maximum_suppression
    g = gold('resnet8_u32_pretrained')
    s, c = non_maximum_suppression(g['y_full'][0, 0], 6, -6.0)
    assert np.array_equal(c, g['nms_coords']) and np.array_equal(s, g['nms_scores'])


def test_fcnn_and_affine_denoisers():
    from topaz_b200.denoising.models import DenoiseNet2, AffineDenoise
    from topaz_b200.denoise import Denoise
    g = gold('fcnn_affine_seeded')
    mf = _load(DenoiseNet2(64, width=11), seeded_state({k: tuple(v.shape) for k, v in DenoiseNet2(64, width=11).state_dict().items()}, 301)).cuda()
    ma = _load(AffineDenoise(max_size=31), seeded_state({k: tuple(v.shape) for k, v in AffineDenoise(31).state_dict().items()}, 302)).cuda()
    x = torch.from_numpy(g['x']).cuda()
    with torch.no_grad():
        _check(mf(x).cpu().numpy(), g['y_fcnn'], TOL_SEEDED)
        _check(ma(x).cpu().numpy(), g['y_affine'], 1e-5)
    img = (10 + 3 * g['x'][0, 0]).astype(np.float32)
    for m, key in ((mf, 'y_fcnn'), (ma, 'y_affine')):
        out = Denoise(m)._denoise(img.copy())
        assert out.shape == img.shape and np.isfinite(out).all()


def test_gpu_nms3d_bit_exact():
    """3-D greedy NMS (algorithms.py:66-103) vs the reference goldens and, on a larger volume, the oracle."""
    from topaz_b200.algorithms import non_maximum_suppression_3d
    g = gold('nms3d')
    for tag in 'abcd':
        s, c = non_maximum_suppression_3d(g[f'{tag}.x'], float(g[f'{tag}.r']), scale=float(g[f'{tag}.scale']), threshold=float(g[f'{tag}.thr']))
        assert s.dtype == np.float32 and c.dtype == np.int32 and c.shape[1] == 3
        assert np.array_equal(c, g[f'{tag}.coords']) and np.array_equal(s, g[f'{tag}.scores']), tag
    x = np.random.default_rng(9).standard_normal((64, 80, 96)).astype(np.float32)
    s_ref, c_ref = O.nms3d(x, 5, 1.0, 1.0)
    s, c = non_maximum_suppression_3d(x, 5, threshold=1.0)
    assert len(s) > 300 and np.array_equal(c, c_ref) and np.array_equal(s, s_ref)


def test_downsample_and_gmm_normalize_gpu():
    """Fourier-crop downsample (utils/image.py:38-61) and GMM normalisation (stats.py:36-214) on the
    GPU vs the reference goldens and, at larger sizes, the oracle."""
    from topaz_b200 import preprocess, stats
    g = gold('preprocess')
    x = g['x']
    for tag, kw in [('f2', dict(factor=2)), ('f4', dict(factor=4)), ('f3', dict(factor=3)), ('s', dict(shape=(37, 50))),
                    ('f1p7', dict(factor=1.7))]:
        y = preprocess.downsample(x, **kw)
        ref = g['ds.' + tag]
        assert y.shape == ref.shape and y.dtype == np.float32
        assert np.abs(y - ref).max() < 1e-4 * np.abs(ref).max(), tag
    big = (100 + 5 * np.random.default_rng(8).standard_normal((1900, 2100))).astype(np.float32)
    for kw in (dict(factor=4), dict(factor=8), dict(shape=(333, 512))):
        y, ref = preprocess.downsample(big, **kw), O.downsample(big, **kw)
        assert y.shape == ref.shape and np.abs(y - ref).max() < 1e-4 * np.abs(ref).max()
    qs = [0.0, 0.02, 0.05, 0.1, 0.5, 0.77, 0.9, 1.0]
    np.testing.assert_allclose(stats.quantiles(torch.from_numpy(big).cuda().view(-1), qs), np.quantile(big.astype(np.float64), qs), rtol=1e-7)
    for tag in 'abc':
        img = g[f'n.{tag}.x']
        y, md = stats.normalize(img, alpha=float(g[f'n.{tag}.alpha']), beta=float(g[f'n.{tag}.beta']),
                                num_iters=int(g[f'n.{tag}.iters']), method='gmm')
        assert abs(md['mu'] - g[f'n.{tag}.mu']) < 1e-4 * abs(md['mu']) and abs(md['std'] - g[f'n.{tag}.std']) < 1e-3 * md['std'], tag
        assert abs(md['pi'] - g[f'n.{tag}.pi']) < 1e-3
        np.testing.assert_allclose(md['logps'], g[f'n.{tag}.logps'], rtol=2e-5)
        assert y.dtype == np.float32 and np.abs(y - g[f'n.{tag}.y']).max() < 1e-3
    # a 480x512 micrograph-like image with a real two-component structure (Beta(2,2) prior so the mixture wins)
    gg = np.random.default_rng(77)
    img = (50 + 3 * gg.standard_normal((480, 512)) + 9 * (gg.random((480, 512)) < 0.2)).astype(np.float32)
    y, md = stats.normalize(img, alpha=2, beta=2, num_iters=60, method='gmm')
    yr, mu, std, pi = O.gmm_normalize(img, 2, 2, 60)
    assert abs(md['mu'] - mu) < 1e-3 * abs(mu) and abs(md['std'] - std) < 1e-3 * std and abs(md['pi'] - pi) < 1e-3
    assert np.abs(y - yr).max() < 1e-3 * np.abs(yr).max()


def test_pick_arrays_matches_score_then_nms():
    """Device-resident score -> NMS pipeline equals score_arrays followed by the oracle NMS on the same score map."""
    from topaz_b200.extract import pick_arrays, score_arrays, nms_iterator
    g = gold('resnet8_u32_pretrained'); sd = weights_of(g)
    m = _load(_classifier('resnet8', 32), sd)
    imgs = [np.random.default_rng(50 + i).standard_normal((300 + 40 * i, 280)).astype(np.float32) for i in range(3)]
    picks = list(pick_arrays(m, imgs, radius=6, threshold=-3.0))
    m.unfill()
    maps = list(score_arrays(m, imgs))
    assert len(picks) == 3
    for (s, c), y in zip(picks, maps):
        s_ref, c_ref = O.nms(y, 6, -3.0)
        assert len(s) > 20 and np.array_equal(c, c_ref) and np.array_equal(s, s_ref)
    out = list(nms_iterator([('a', maps[0])], 6, -3.0))
    assert out[0][0] == 'a' and np.array_equal(out[0][2], picks[0][1])


@pytest.mark.parametrize('k,co,slope', [(7, 64, 0.0), (7, 32, 0.0), (11, 48, 0.1), (5, 20, 1.0), (3, 64, 0.25), (11, 32, 0.0)])
def test_fused_first_layer_kernel_matches_fp32_conv(k, co, slope):
    """tpz_conv_first_tc (im2col tile built in smem + tcgen05) vs a plain fp32 torch convolution of the same Cin=1
    layer, on ragged sizes and batch > 1; tolerance = fp16 operand rounding (2^-11) over k*k taps."""
    from topaz_b200 import ops
    g = torch.Generator().manual_seed(k * 100 + co)
    w = torch.randn(co, k, k, generator=g) / k
    b = torch.randn(co, generator=g)
    cp = (co + 31) // 32 * 32
    wp, bp = ops.pack_first_tc(w, b, cp, 'cuda')
    for (B, H, W, pad) in [(1, 64, 64, k // 2), (2, 37, 53, 35 if k == 7 else k // 2), (1, 5, 200, k // 2), (1, 300, 9, k // 2)]:
        x = torch.randn(B, H, W, generator=g)
        y = ops.conv_first_tc(x.cuda(), wp, bp, k, pad, slope).float().cpu()          # [B,1,Ho,Wo,cp]
        ref = F.conv2d(x[:, None], w[:, None], b, padding=pad)
        ref = torch.where(ref > 0, ref, ref * slope).permute(0, 2, 3, 1)
        assert y.shape[:4] == (B, 1) + tuple(ref.shape[1:3])
        assert torch.all(y[..., co:] == 0)                                             # padded channels stay zero
        err = (y[:, 0, :, :, :co] - ref).abs().max() / ref.abs().max()
        assert err < 2e-3, (B, H, W, float(err))


@pytest.mark.parametrize('H,W', [(64, 64), (37, 53), (130, 18), (9, 200)])
def test_fused_first_layer_with_pool_matches_unfused(H, W):
    """tpz_conv_first_tc(pool=1) == maxpool2(tpz_conv_first_tc(pool=0)) bit for bit (odd sizes: floor, like MaxPool2d)."""
    from topaz_b200 import ops
    g = torch.Generator().manual_seed(H * 7 + W)
    w = torch.randn(48, 11, 11, generator=g) / 11
    b = torch.randn(48, generator=g)
    wp, bp = ops.pack_first_tc(w, b, 64, 'cuda')
    x = torch.randn(2, H, W, generator=g).cuda()
    full = ops.conv_first_tc(x, wp, bp, 11, 5, 0.1)
    ref = ops.maxpool2(full, 2)
    got = ops.conv_first_tc(x, wp, bp, 11, 5, 0.1, pool=True)
    assert got.shape == ref.shape and torch.equal(got, ref)
