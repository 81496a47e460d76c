"""GPU drop-in for topaz.algorithms.non_maximum_suppression (reference algorithms.py:25-63): greedy NMS over a score
map, returning (scores float32 [j], coords int32 [j,2] as (x, y)) in descending score order.  Picks are bit-identical to
the reference's sequential loop, including its clip-to-shape quirk at the right border; ties (equal scores) are ordered by
larger flat index first (the reference's unstable argsort leaves tie order unspecified)."""
import ctypes as C

import numpy as np
import torch

from topaz_b200 import _lib, ops
from topaz_b200._lib import check


def non_maximum_suppression(x, r: int, threshold: float = -np.inf, max_picks: int = None):
    xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda() if isinstance(x, np.ndarray) else x.contiguous().float()
    ops.require_cuda(xd, 'score map')
    H, W = xd.shape
    if max_picks is None:
        max_picks = min(H * W, max(1024, (H * W) // max(1, (r * r) // 2 + 1) + 1024))
    state = torch.empty(H * W, dtype=torch.uint8, device=xd.device)
    lst = torch.empty(max_picks, dtype=torch.int32, device=xd.device)
    counters = torch.empty(2, dtype=torch.int32, device=xd.device)
    n = C.c_int(0)
    thr = float(threshold) if np.isfinite(threshold) else (-3.4e38 if threshold < 0 else 3.4e38)
    ops._count(4)
    check(_lib.lib().tpz_nms2d(C.c_void_p(xd.data_ptr()), H, W, int(r), thr, C.c_void_p(state.data_ptr()),
                               C.c_void_p(lst.data_ptr()), C.c_void_p(counters.data_ptr()), max_picks, C.byref(n),
                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    idx = lst[:n.value].cpu().numpy().astype(np.int64)
    sc = xd.reshape(-1)[lst[:n.value].long()].cpu().numpy().astype(np.float32)
    order = np.lexsort((-idx, -sc.astype(np.float64)))          # score descending, ties: larger flat index first
    idx, sc = idx[order], sc[order]
    coords = np.stack([idx % W, idx // W], axis=1).astype(np.int32) if len(idx) else np.zeros((0, 2), dtype=np.int32)
    return sc, coords
