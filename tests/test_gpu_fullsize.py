"""-m gpu parity at BASELINE.json's FULL sizes (4096x4096 micrographs), where the CPU oracle cannot run the whole image in
test time.  Each case checks the full-size CUDA result against the oracle on what the oracle CAN do in seconds (windows
whose receptive field is cut from the same input; one whole patch of the patched denoiser; the complete greedy NMS) plus
a size-independent property of the domain (translation equivariance of the dense classifier)."""
import numpy as np
import pytest
import torch

from common import gold, weights_of, rel_err, rel_err_dc, check_parity
from oracle import topaz_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _classifier(units):
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    return LinearClassifier(get_feature_extractor('resnet8', units=units, bn=False))


def test_resnet8_u64_4096_windows_vs_oracle_and_translation_equivariance():
    """configs[1]: ResNet8-u64 over one 4096^2 micrograph (extract.py:224-256 -> classifier.py:48-66).  Output pixel
    (i, j) of the filled network depends only on input[i-35:i+36, j-35:j+36] (zero outside), so the oracle run on a
    window + 35-px ring reproduces that window of the full-size result."""
    from topaz_b200.extract import score_arrays
    g = gold('resnet8_u64_pretrained'); sd = weights_of(g)
    m = _classifier(64)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})      # score_arrays fills the model itself, like
    S = 4096                                                                   # extract.score_images (extract.py:231-232)
    x = np.random.default_rng(1000).standard_normal((S, S)).astype(np.float32)
    (y,) = list(score_arrays(m, [x]))
    assert y.shape == (S, S) and y.dtype == np.float32 and np.isfinite(y).all()
    scale = np.abs(y).max()
    halo, win = 35, 96
    rng = np.random.default_rng(5)
    corners = [(0, 0), (S - win, S - win), (0, S - win), (2000, 2040)] + [tuple(rng.integers(0, S - win, 2)) for _ in range(2)]
    xp = np.pad(x, halo)
    for (i, j) in corners:
        crop = xp[i:i + win + 2 * halo, j:j + win + 2 * halo]                  # window + receptive-field ring
        # the oracle pads by 35 itself (resnet.py:240-243): feed the crop and keep its centre
        ref = O.classifier_forward(sd, crop[None, None], 'resnet8', 64, filled=True).numpy()[0, 0][halo:halo + win, halo:halo + win]
        got = y[i:i + win, j:j + win]
        assert np.abs(got - ref).max() / scale < TOL, (i, j, np.abs(got - ref).max() / scale)
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < TOL
    # translation equivariance: scoring the image shifted by (37, 101) gives the shifted scores away from the borders
    # (different tile decomposition, same arithmetic per output pixel)
    di, dj = 37, 101
    m.unfill()
    (y2,) = list(score_arrays(m, [np.ascontiguousarray(x[di:, dj:])]))
    a = y[di + halo:S - halo, dj + halo:S - halo]
    b = y2[halo:S - di - halo, halo:S - dj - halo]
    assert np.abs(a - b).max() / scale < 1e-5


def test_unet_4096_patched_denoise_one_patch_vs_oracle():
    """configs[2]: Denoise.denoise(4096^2, patch_size=1024, padding=500) (denoise.py:299-332).  Patch (0, 0) is the
    network applied to x[0:1524, 0:1524] normalised by that crop's own mean / unbiased std; the oracle runs that one
    crop.  The remaining patches are checked for seam-free finiteness and against the device-resident path."""
    from topaz_b200.denoising.models import UDenoiseNet
    from topaz_b200.denoise import Denoise
    sd = weights_of(gold('unet_pretrained'))
    m = UDenoiseNet(base_width=11, top_width=5)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    dn = Denoise(m)
    S, ps, pad = 4096, 1024, 500
    x = (10 + 3 * np.random.default_rng(3000).standard_normal((S, S))).astype(np.float32)
    y = dn.denoise(x, patch_size=ps, padding=pad)
    assert y.shape == (S, S) and np.isfinite(y).all()
    ref = O.denoise_call(sd, x[:ps + pad, :ps + pad])[:ps, :ps]
    got = y[:ps, :ps]
    # the denoised raw-like image is mean 10 / std 0.1: the DC-free metric holds the error to 1e-3 of the denoised SIGNAL
    check_parity(got, ref, TOL, 'cfg3 patch (0,0) of the 4096^2 patched denoise', dc_free=True)
    yd = dn.denoise_patches_device(torch.from_numpy(x).cuda(), ps, pad).cpu().numpy()
    assert np.array_equal(yd, y)


def test_nms_4096_bit_exact_vs_oracle():
    """a17 at full size: greedy NMS over a 4096^2 map with 2^24 DISTINCT scores (so the reference's unspecified tie order
    cannot matter), r = 8, threshold 1.6: identical picks, scores and order as the sequential reference loop."""
    from topaz_b200.algorithms import non_maximum_suppression
    S = 4096
    vals = ((np.arange(S * S, dtype=np.int64) - (S * S) // 2).astype(np.float32)) * np.float32(2.0 ** -22)   # exact, distinct
    x = vals[np.random.default_rng(11).permutation(S * S)].reshape(S, S)
    r, thr = 8, 1.6
    s, c = non_maximum_suppression(x, r, thr)
    s_ref, c_ref = O.nms(x, r, thr)
    assert len(s) == len(s_ref) and len(s) > 10000
    assert np.array_equal(c, c_ref) and np.array_equal(s, s_ref)
    # properties: descending, above threshold, no pick inside an earlier pick's disc
    assert np.all(np.diff(s) < 0) and s.min() > thr
    from scipy.spatial import cKDTree
    pairs = cKDTree(c.astype(np.float64)).query_pairs(r)
    assert len(pairs) == 0
