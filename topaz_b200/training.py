"""Drop-in for the data side of topaz.training: ``make_data_iterators`` (reference training.py:479-503) with the training
iterator backed by the GPU crop sampler (topaz_b200.sampler.GpuCropSampler, csrc/tpz_sampler.cu) instead of
``MultipleImageSetDataset`` + ``DataLoader`` (0.28-0.35 s of host work per 256-crop minibatch against a ~2.4 ms step).

``compat.install()`` patches it into ``topaz.training``, so ``topaz train`` / ``python -m topaz_b200 train`` pick it up with
no change to the reference's command code.  Coordinate parsing, path grouping, disc expansion and the TEST iterator are the
reference's own functions (file formats are outside the hot path); only the per-minibatch sampling moved to the device.
Reproducibility: minibatch b of a run is a function of (seed, b) (Philox), seed = torch.initial_seed()."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch


class GpuCropIterator:
    """What ``fit_epoch`` iterates (training.py:551-568): ``epoch_size`` minibatches ``(X [B, crop, crop] float32 cuda,
    Y [B] float64 cuda)`` per pass; ``len()`` = epoch_size; ``batch_size`` like a DataLoader."""

    def __init__(self, sampler, minibatch_size: int, epoch_size: int):
        self.sampler, self.batch_size, self.epoch_size = sampler, int(minibatch_size), int(epoch_size)
        self.dataset = sampler          # `.dataset` as on a DataLoader (the reference reports counts from it)

    def __len__(self):
        return self.epoch_size

    def __iter__(self):
        for _ in range(self.epoch_size):
            yield self.sampler.sample(self.batch_size)


def _first(v):
    return v[0] if isinstance(v, tuple) else v


def expand_target_points(targets, radius: int):
    """Particle centres -> every pixel of the disc of `radius` around them (reference training.py:444-476, 2-D):
    returns (DataFrame[image_name, x_coord, y_coord], pixels per disc)."""
    r = int(np.floor(radius))
    yy, xx = np.mgrid[-r:r + 1, -r:r + 1]
    keep = (xx ** 2 + yy ** 2) <= radius ** 2
    import pandas as pd
    offs = pd.DataFrame({'y_offset': yy[keep], 'x_offset': xx[keep]})
    e = targets[['image_name', 'x_coord', 'y_coord']].merge(offs, how='cross')
    e['x_coord'] = e['x_coord'] + e['x_offset']
    e['y_coord'] = e['y_coord'] + e['y_offset']
    return e[['image_name', 'x_coord', 'y_coord']], int(keep.sum())


def gpu_training_iterator(image_groups, expanded_targets, crop: int, split, minibatch_size: int, epoch_size: int,
                          balance: float = 0.5, load=None, seed=None) -> GpuCropIterator:
    """The training iterator of make_data_iterators on the GPU sampler.  image_groups: list (one entry per source) of lists
    of micrograph paths; expanded_targets: DataFrame[image_name, x_coord, y_coord] with EVERY labelled pixel (discs already
    expanded); load(path) -> 2-D float array (default: topaz_b200.mrc.read)."""
    from topaz_b200.sampler import GpuCropSampler
    if load is None:
        from topaz_b200 import mrc
        load = lambda p: _first(mrc.read(p))
    t = expanded_targets.copy()
    t[['y_coord', 'x_coord']] = t[['y_coord', 'x_coord']].round().astype(int)                     # memory_mapped_data.py:136
    image_sets, index_of = [], {}
    for group in image_groups:
        imgs = []
        for path in group:
            name = os.path.splitext(path.split('/')[-1])[0]
            index_of[name] = sum(len(s) for s in image_sets) + len(imgs)
            imgs.append(np.asarray(load(path), dtype=np.float32))
        image_sets.append(imgs)
    idx = t['image_name'].map(index_of)
    known = idx.notna().to_numpy()
    pos = np.stack([idx[known].to_numpy().astype(np.int64), t['y_coord'].to_numpy()[known], t['x_coord'].to_numpy()[known]],
                   axis=1) if known.any() else np.zeros((0, 3), dtype=np.int64)
    seed = int(torch.initial_seed() % (1 << 62)) if seed is None else int(seed)
    sampler = GpuCropSampler(image_sets, pos, crop, positive_balance=balance, split=split, rotate=True, flip=True, seed=seed)
    return GpuCropIterator(sampler, minibatch_size, epoch_size)


def make_data_iterators(train_image_path: str, train_targets_path: str, crop: int, split, minibatch_size: int, epoch_size: int,
                        test_image_path: str = None, test_targets_path: str = None, testing_batch_size: int = 1,
                        num_workers: int = 0, balance: float = 0.5, dims: int = 2, use_cuda: bool = False, radius: int = 3):
    '''make train and test iterators (reference training.py:479-503); 2-D training data is sampled on the GPU'''
    import topaz.training as ref
    orig = getattr(ref, '_tpz_orig_make_data_iterators', None)
    if dims != 2 or not torch.cuda.is_available() or os.environ.get('TPZ_GPU_SAMPLER', '1') == '0':
        if orig is None:
            raise RuntimeError('topaz_b200: the GPU crop sampler covers 2-D training on a CUDA device only')
        return orig(train_image_path, train_targets_path, crop, split, minibatch_size, epoch_size, test_image_path,
                    test_targets_path, testing_batch_size, num_workers, balance, dims, use_cuda, radius)
    from topaz.utils.data.loader import load_image
    train_targets = ref.file_utils.read_coordinates(train_targets_path)
    if len(train_targets) == 0:
        ref.report('ERROR: no training particles specified. Check that micrograph names in the particles file match those in the micrographs file/directory.', file=sys.stderr)
        raise Exception('No training particles.')
    groups = ref.convert_path_to_grouped_list(train_image_path, train_targets)
    expanded, mask_size = ref.expand_target_points(train_targets, radius, dims)
    train_iter = gpu_training_iterator(groups, expanded, crop, split, minibatch_size, epoch_size, balance,
                                       load=lambda p: load_image(p, make_image=False, return_header=False))
    n_img = sum(len(g) for g in groups)
    ref.report(f'Loaded {n_img} training micrographs with ~{int(train_iter.sampler.positives.shape[0] // max(1, int(mask_size)))} labeled particles '
               f'(crops are sampled on the GPU)')
    test_iter = None
    if test_targets_path is not None:
        from torch.utils.data import DataLoader
        test_targets = ref.file_utils.read_coordinates(test_targets_path)
        test_dataset = ref.TestingImageDataset(test_image_path, test_targets, radius=radius, dims=dims, use_cuda=use_cuda)
        test_iter = DataLoader(test_dataset, batch_size=testing_batch_size, shuffle=False, num_workers=num_workers)
        ref.report(f'Loaded {len(test_dataset)} testing micrographs with {len(test_targets)} labeled particles')
    return train_iter, test_iter
