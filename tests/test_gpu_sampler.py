"""-m gpu tests of the GPU crop sampler: (1) crop / rotate / flip geometry bit-identical to the reference pipeline
(numpy zero-padded crop -> torchvision rotate(NEAREST) -> centre crop -> hflip / vflip, memory_mapped_data.py:45-70,
215-231) for explicit parameters; (2) sampling statistics: positive fraction, labels, 'pn' rejection of labelled pixels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference_crop(img, cy, cx, crop, angle, hflip, vflip):
    import torchvision.transforms.functional as TF
    big = int(np.ceil(crop * np.sqrt(2)))
    if (big - crop) % 2:
        big += 1
    H, W = img.shape
    xmin, xmax, ymin, ymax = cx - big // 2, cx + big // 2 + 1, cy - big // 2, cy + big // 2 + 1
    xpad = abs(min(0, xmin)), abs(min(0, W - xmax)); ypad = abs(min(0, ymin)), abs(min(0, H - ymax))
    c = np.pad(img[max(0, ymin):ymax, max(0, xmin):xmax], (ypad, xpad))
    t = torch.from_numpy(c)[None]
    t = TF.rotate(t, float(angle))
    d = (t.shape[-1] - crop) // 2
    t = t[..., d:d + crop, d:d + crop]
    if hflip: t = TF.hflip(t)
    if vflip: t = TF.vflip(t)
    return t[0].numpy()


def test_crop_rotate_flip_geometry_matches_torchvision_pipeline():
    from topaz_b200.sampler import GpuCropSampler
    rng = np.random.default_rng(0)
    imgs = [rng.standard_normal((180, 200)).astype(np.float32), rng.standard_normal((150, 130)).astype(np.float32)]
    s = GpuCropSampler([imgs], np.zeros((0, 3), dtype=np.int32), 71)
    assert s.big == 101
    params = [(0, 90, 100, 1, 0.0, 0, 0), (0, 5, 3, 0, 37.25, 1, 0), (1, 149, 129, 0, 123.5, 0, 1), (1, 70, 60, 1, 289.75, 1, 1),
              (0, 100, 199, 0, 90.0, 0, 0), (1, 0, 64, 0, 180.0, 1, 0)]
    X = s.crops_for(np.array(params, dtype=np.float64)).cpu().numpy()
    total = 0
    for b, (i, cy, cx, _, ang, hf, vf) in enumerate(params):
        ref = _reference_crop(imgs[int(i)], int(cy), int(cx), 71, ang, hf, vf)
        bad = int((X[b] != ref).sum()); total += bad
        assert bad <= 2, (b, bad)            # nearest-neighbour ties within 1 ulp of x.5 may fall either way
    assert total <= 4


def test_sampling_statistics_and_labels():
    from topaz_b200.sampler import GpuCropSampler
    rng = np.random.default_rng(1)
    imgs = [np.arange(120 * 140, dtype=np.float32).reshape(120, 140) + 1e6 * k for k in range(3)]   # pixel value encodes location
    pos = []
    for k in range(3):
        for (py, px) in [(30, 40), (80, 100)]:
            for dy in range(-3, 4):
                for dx in range(-3, 4):
                    if dy * dy + dx * dx <= 9:
                        pos.append((k, py + dy, px + dx))
    pos = np.array(pos, dtype=np.int32)
    s = GpuCropSampler([[imgs[0], imgs[1]], [imgs[2]]], pos, 71, image_set_balance=[0.25, 0.75], positive_balance=0.0625,
                       rotate=False, flip=False, seed=3)
    B = 20000
    X, Y = s.sample(B)
    Y = Y.cpu().numpy(); X = X.cpu().numpy(); prm = s.last_params.cpu().numpy()
    assert Y.dtype == np.float64 and abs(Y.mean() - 0.0625) < 0.01
    centre = X[:, 35, 35]
    img_of = (centre // 1e6).astype(int); loc = (centre % 1e6).astype(int)
    cy, cx = loc // 140, loc % 140
    assert np.array_equal(img_of, prm[:, 0]) and np.array_equal(cy, prm[:, 1]) and np.array_equal(cx, prm[:, 2])
    posset = {tuple(p) for p in pos.tolist()}
    is_pos = np.array([(int(a), int(b), int(c)) in posset for a, b, c in zip(img_of, cy, cx)])
    assert np.array_equal(is_pos, Y == 1)                       # positives land on labelled pixels, negatives never do
    neg = Y == 0
    frac_set1 = (img_of[neg] == 2).mean()
    assert abs(frac_set1 - 0.75) < 0.02                         # image_set_balance respected
    X2, Y2 = s.sample(B)
    assert not np.array_equal(Y2.cpu().numpy(), Y)              # a new batch index draws new samples
    s2 = GpuCropSampler([[imgs[0], imgs[1]], [imgs[2]]], pos, 71, image_set_balance=[0.25, 0.75], positive_balance=0.0625,
                        rotate=False, flip=False, seed=3)
    X3, Y3 = s2.sample(B)
    assert torch.equal(Y3.cpu(), torch.from_numpy(Y)) and np.array_equal(X3.cpu().numpy(), X)   # reproducible from the seed


def test_gpu_training_iterator_replaces_the_reference_data_loader(tmp_path):
    """topaz_b200.training.gpu_training_iterator (what the make_data_iterators drop-in builds, reference training.py:479-503):
    micrographs on disk + a particle table -> epoch_size minibatches (X [B, crop, crop] float32, Y [B] float64) on the device;
    positives are drawn from the expanded discs, their crops are centred on labelled pixels, and a GE_binomial epoch runs on it."""
    import pandas as pd
    import torch.nn as nn
    from topaz_b200 import mrc
    from topaz_b200.training import gpu_training_iterator, expand_target_points
    rng = np.random.default_rng(5)
    paths, rows = [], []
    for k in range(3):
        img = (rng.standard_normal((300, 320)) * 0.1).astype(np.float32)
        for (py, px) in rng.integers(40, 260, size=(6, 2)):
            img[py - 3:py + 4, px - 3:px + 4] = 5.0                  # a bright blob under every particle's radius-3 disc
            rows.append((f'mic{k}', int(px), int(py)))
        p = str(tmp_path / f'mic{k}.mrc'); mrc.write(p, img); paths.append(p)
    targets = pd.DataFrame(rows, columns=['image_name', 'x_coord', 'y_coord'])
    expanded, mask_size = expand_target_points(targets, 3)
    assert mask_size == 29 and len(expanded) == 29 * len(targets)
    it = gpu_training_iterator([paths[:2], paths[2:]], expanded, 71, 'pn', 256, 5, balance=0.25, seed=11)
    assert len(it) == 5 and it.batch_size == 256
    frac, n = 0.0, 0
    for X, Y in it:
        assert X.is_cuda and X.dtype == torch.float32 and tuple(X.shape) == (256, 71, 71)
        assert Y.is_cuda and Y.dtype == torch.float64 and tuple(Y.shape) == (256,)
        centre = X[:, 35, 35]
        pos = Y == 1
        assert bool((centre[pos] > 2.5).all())                      # positives sit on a blob (rotation / flips keep the centre)
        assert float((centre[~pos] > 2.5).float().mean()) < 0.05    # 'pn': unlabeled crops avoid labelled pixels
        frac += float(pos.float().mean()); n += 1
    assert abs(frac / n - 0.25) < 0.06
    # one epoch of the reference's fit loop shape: step() on every minibatch
    from topaz_b200.methods import GE_binomial
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    m = LinearClassifier(get_feature_extractor('resnet8', units=32, bn=False)).cuda(); m.train()
    tr = GE_binomial(m, torch.optim.Adam(m.parameters(), lr=2e-4), nn.BCEWithLogitsLoss(), 0.035)
    outs = [tr.step(X, Y.view(-1)) for X, Y in it]
    assert len(outs) == 5 and all(np.isfinite(o).all() for o in outs)
