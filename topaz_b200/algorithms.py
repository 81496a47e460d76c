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


def non_maximum_suppression_3d(x, r, scale: float = 1.0, threshold: float = -np.inf, max_picks: int = None):
    """GPU drop-in for topaz.algorithms.non_maximum_suppression_3d (reference algorithms.py:66-103): returns
    (scores float32 [j], coords int32 [j,3] as (x, y, z)).  The reference suppresses flat indices i + delta without
    bounds handling (neighbourhoods wrap across rows and slices); that behaviour is reproduced exactly."""
    xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda() if isinstance(x, np.ndarray) else x.contiguous().float()
    ops.require_cuda(xd, 'score volume')
    D, H, W = xd.shape
    rr = scale * r
    width = int(np.ceil(rr))
    ax = np.arange(-width, width + 1)
    ii, jj, kk = np.meshgrid(ax, ax, ax)
    inside = (ii ** 2 + jj ** 2 + kk ** 2) <= rr * rr
    deltas = np.unique((ii[inside] * (H * W) + jj[inside] * W + kk[inside]).astype(np.int64))
    n = D * H * W
    if n >= 2 ** 31 or np.abs(deltas).max(initial=0) >= 2 ** 31:
        raise ValueError('volume too large for 32-bit flat indices')
    dd = torch.from_numpy(deltas.astype(np.int32)).to(xd.device)
    if max_picks is None:
        max_picks = min(n, max(1024, n // max(1, len(deltas) // 8) + 1024))
    state = torch.empty(n, dtype=torch.uint8, device=xd.device)
    lst = torch.empty(max_picks, dtype=torch.int32, device=xd.device)
    counters = torch.empty(2, dtype=torch.int32, device=xd.device)
    cnt = C.c_int(0)
    thr = float(threshold) if np.isfinite(threshold) else (-3.4e38 if threshold < 0 else 3.4e38)
    ops._count(4)
    check(_lib.lib().tpz_nms_flat(C.c_void_p(xd.data_ptr()), n, C.c_void_p(dd.data_ptr()), len(deltas), thr,
                                  C.c_void_p(state.data_ptr()), C.c_void_p(lst.data_ptr()), C.c_void_p(counters.data_ptr()),
                                  max_picks, C.byref(cnt), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    idx = lst[:cnt.value].cpu().numpy().astype(np.int64)
    sc = xd.reshape(-1)[lst[:cnt.value].long()].cpu().numpy().astype(np.float32)
    order = np.lexsort((-idx, -sc.astype(np.float64)))
    idx, sc = idx[order], sc[order]
    zz, yy, xx = np.unravel_index(idx, (D, H, W)) if len(idx) else (np.zeros(0, int),) * 3
    return sc, np.stack([xx, yy, zz], axis=1).astype(np.int32)
