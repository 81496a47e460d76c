"""GPU-resident training data: positive-balanced crop sampling + rotate/flip augmentation in two kernel launches per
minibatch (csrc/tpz_sampler.cu).  Drop-in for the role of ``MultipleImageSetDataset`` + ``DataLoader`` in
``topaz.training.make_data_iterators`` (training.py:479-503): iterate it to get ``(X [B,crop,crop] float32 cuda,
Y [B] float64 cuda)`` minibatches for ``GE_binomial.step``.  In the reference this host loop costs 0.28-0.35 s per
256-crop minibatch (per-sample file open + memmap + pandas .sample() + torchvision rotate) against a ~3 ms step."""
import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from topaz_b200 import _lib, ops
from topaz_b200._lib import check


class GpuCropSampler:
    def __init__(self, image_sets: Sequence[Sequence[np.ndarray]], positives: np.ndarray, crop_size: int,
                 image_set_balance: Optional[Sequence[float]] = None, positive_balance: float = 0.5, split: str = 'pn',
                 rotate: bool = True, flip: bool = True, seed: int = 0, device='cuda'):
        """image_sets: list of sets, each a list of 2-D float32 micrographs; positives: int array [P,3] of (global image
        index, y, x) for EVERY labelled pixel (the reference expands particle centres to discs, training.py:492)."""
        self.device = torch.device(device)
        self.crop = int(crop_size)
        self.big = int(np.ceil(crop_size * np.sqrt(2))) if rotate else int(crop_size)   # memory_mapped_data.py:144
        # keep (big - crop) even so that the final crop is centred on the rotation centre.  Model widths are odd (ResNet8: 71,
        # conv31/63/127), for which this is exactly the reference's geometry (2*(size//2)+1 window); an EVEN crop size would put the
        # rotation centre half a pixel from the reference's (statistically equivalent augmentation, not the same pixels).
        if (self.big - self.crop) % 2:
            self.big += 1
        flat, recs, set_begin, off = [], [], [0], 0
        for s in image_sets:
            for im in s:
                a = np.ascontiguousarray(im, dtype=np.float32)
                recs.append((off, a.shape[0], a.shape[1])); flat.append(a.reshape(-1)); off += a.size
            set_begin.append(len(recs))
        self.shapes = [(h, w) for _, h, w in recs]
        self.pixels = torch.from_numpy(np.concatenate(flat)).to(self.device)
        rec = np.zeros(len(recs), dtype=[('offset', '<i8'), ('H', '<i4'), ('W', '<i4')])
        for i, r in enumerate(recs):
            rec[i] = r
        self.imgs = torch.from_numpy(rec.view(np.uint8).copy()).to(self.device)
        self.set_begin = torch.tensor(set_begin, dtype=torch.int32, device=self.device)
        n = len(image_sets)
        bal = np.full(n, 1.0 / n) if image_set_balance is None else np.asarray(image_set_balance, dtype=np.float64)
        self.set_cdf = torch.tensor(np.cumsum(bal / bal.sum()), dtype=torch.float32, device=self.device)
        self.nsets = n
        pos = np.ascontiguousarray(positives, dtype=np.int32).reshape(-1, 3)
        ok = np.array([0 <= y < recs[i][1] and 0 <= x < recs[i][2] for i, y, x in pos], dtype=bool) if len(pos) else np.zeros(0, bool)
        pos = pos[ok]                                   # out-of-bounds labels are dropped (memory_mapped_data.py:101-112)
        self.positives = torch.from_numpy(pos.copy()).to(self.device)
        mask = np.zeros(off, dtype=np.uint8)
        for i, y, x in pos:
            mask[recs[i][0] + y * recs[i][2] + x] = 1
        self.pos_mask = torch.from_numpy(mask).to(self.device)
        self.positive_balance, self.split_pn = float(positive_balance), int(split == 'pn' and len(pos) > 0)
        self.rotate, self.flip, self.seed = int(rotate), int(flip), int(seed)
        self.batch_index = 0

    def sample(self, B: int):
        X = torch.empty((B, self.crop, self.crop), dtype=torch.float32, device=self.device)
        Y = torch.empty(B, dtype=torch.float64, device=self.device)
        scratch = torch.empty(B * 8, dtype=torch.int32, device=self.device)
        P = lambda t: C.c_void_p(t.data_ptr())
        ops._count(2)
        check(_lib.lib().tpz_sample_crops(B, self.seed, self.batch_index, P(self.imgs), P(self.pixels), P(self.set_begin),
                                          P(self.set_cdf), self.nsets, P(self.positives), int(self.positives.shape[0]),
                                          P(self.pos_mask), self.positive_balance, self.split_pn, self.rotate, self.flip,
                                          self.crop, self.big, P(scratch), P(X), P(Y),
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self.batch_index += 1
        self.last_params = scratch.view(B, 8)
        return X, Y

    def crops_for(self, params: np.ndarray) -> torch.Tensor:
        """Deterministic crops for explicit per-sample parameters [B,7] = (img, cy, cx, label, angle_deg, hflip, vflip)."""
        B = len(params)
        rec = np.zeros(B, dtype=[('img', '<i4'), ('cy', '<i4'), ('cx', '<i4'), ('label', '<i4'), ('angle', '<f4'),
                                 ('hflip', '<i4'), ('vflip', '<i4'), ('pad', '<i4')])
        for b, p in enumerate(params):
            rec[b] = (int(p[0]), int(p[1]), int(p[2]), int(p[3]), float(p[4]), int(p[5]), int(p[6]), 0)
        pd = torch.from_numpy(rec.view(np.uint8).copy()).to(self.device)
        X = torch.empty((B, self.crop, self.crop), dtype=torch.float32, device=self.device)
        ops._count(1)
        check(_lib.lib().tpz_make_crops(B, self.crop, self.big, C.c_void_p(self.imgs.data_ptr()), C.c_void_p(self.pixels.data_ptr()),
                                        C.c_void_p(pd.data_ptr()), C.c_void_p(X.data_ptr()),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return X

    def batches(self, minibatch_size: int, num_batches: int):
        for _ in range(num_batches):
            yield self.sample(minibatch_size)
