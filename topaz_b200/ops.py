"""Thin Python wrappers over the C ABI.  torch tensors are containers only (device memory + streams);
every arithmetic op on the hot path is a kernel from libtopaz_b200.so.  Activations are channels-last fp16
tensors of shape [N, D, H, W, C] (2-D: D == 1)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import TcKBlock, TpzTcConvArgs, check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f'topaz_b200: {what} must live on a CUDA device; this build has no CPU path')


# ------------------------------------------------------------------------------------------------
# tensor-core conv plans
# ------------------------------------------------------------------------------------------------
@dataclass
class ConvPart:
    """One input source of a conv: weights [Co, Ci_real, kd, kh, kw] (fp32 torch, any device)."""
    w: torch.Tensor
    c_store: int                 # stored channels of the source tensor (>= Ci_real, multiple of KC)
    dil: int = 1
    org: Tuple[int, int, int] = (0, 0, 0)   # (x, y, z) offset of tap 0 relative to the output pixel
    lat: int = 0                 # lattice spacing of this source in its own pixels (0 = the plan's output lattice)
    lat_z: int = 0               # same along z (0 = the plan's lattice_z)
    phase: bool = True           # False: the output phase does not shift this source (half-resolution source)
    split: bool = False          # the source tensor stores (hi, lo) fp16 pairs: channels [0, c_store) = hi, [c_store, 2*c_store) = lo


@dataclass
class TcConvPlan:
    KC: int
    Co: int                      # stored output channels (multiple of 16)
    kblocks: List[Tuple[int, int, int, int, int]]   # (dx, dy, dz, c0, src)
    orgs: List[Tuple[int, int, int]]
    c_stores: List[int]
    tapgrids: List[Tuple[int, int]]   # (kw, kh) per source
    lattice: int                      # output lattice spacing (in-plane dilation shared by all multi-tap sources)
    weights: torch.Tensor        # [nkb, Co, KC] fp16 (device)
    bias: torch.Tensor           # [Co] fp32 (device)
    neg_slope: float
    lats: List[int] = field(default_factory=list)        # per-source lattice spacing (0 = lattice)
    phases: List[bool] = field(default_factory=list)     # per-source: output phase shifts the source
    phase_sel: int = 0                # 0 = all output phases, k = only phase k-1
    lat_zs: List[int] = field(default_factory=list)
    lattice_z: int = 1                # output z lattice and the z phase this plan computes
    phase_z: int = 0
    dot_w: Optional[torch.Tensor] = None
    dot_b: float = 0.0
    res_scale: Optional[torch.Tensor] = None
    TW: int = 16
    TH: int = 8
    oscale: Optional[torch.Tensor] = None   # [Co] fp32 power-of-two row factors taken out of the fp16 weights (range guard)
    split_out: bool = False           # the output is written as (hi, lo) fp16 pairs: [.., 2*Co] = hi | lo halves
    w_split: bool = False             # weights carried as (hi, lo) fp16 pairs (extra k-blocks)

    @property
    def out_channels(self) -> int:
        """stored channels of this conv's output tensor (split output: hi and lo halves)"""
        return 2 * self.Co if self.split_out else self.Co


def _row_scales(ws: Sequence[torch.Tensor], co_store: int) -> Optional[torch.Tensor]:
    """Range guard for the fp16 weights (BN-folded rows can be huge when running_var is tiny, or tiny when gamma is):
    a row whose largest |w| would overflow fp16 or fall towards its subnormals is divided by 2^floor(log2 max|w|) and the
    factor goes into the fp32 epilogue (TpzTcConvArgs.oscale) -- exact, powers of two.  None when every row is in range."""
    mx = torch.zeros(co_store, dtype=torch.float32)
    for w in ws:
        mx[:w.shape[0]] = torch.maximum(mx[:w.shape[0]], w.reshape(w.shape[0], -1).abs().amax(dim=1))
    if not bool(torch.isfinite(mx).all()):
        raise RuntimeError('topaz_b200: non-finite convolution weights (after BatchNorm folding)')
    bad = (mx > 2.0 ** 14) | ((mx > 0) & (mx < 2.0 ** -10))
    if not bool(bad.any()):
        return None
    sc = torch.ones(co_store, dtype=torch.float32)
    sc[bad] = torch.exp2(torch.floor(torch.log2(mx[bad])))
    return sc


def pack_tc_conv(parts: Sequence[ConvPart], bias: Optional[torch.Tensor], co_store: int, neg_slope: float,
                 device, KC: Optional[int] = None, dot_w=None, dot_b: float = 0.0, res_scale=None,
                 out_scale: Optional[torch.Tensor] = None, lattice: Optional[int] = None, phase_sel: int = 0,
                 lattice_z: int = 1, phase_z: int = 0, strict: bool = False, split_out: Optional[bool] = None) -> TcConvPlan:
    """Repack OIHW fp32 weights into the kernel's [k-block][Co][KC] fp16 layout.

    k-blocks are ordered (source, tap, chunk); all-zero blocks (channel padding) are dropped.
    ``out_scale`` ([Co_real]) multiplies the weight rows (BN eval-mode folding).

    ``strict``: 22-bit operands from fp16 pairs.  The weights are split w = w_hi + w_lo (w_lo = fp16(w - w_hi)); a source
    with ``ConvPart.split`` stores [hi | lo] halves of ``c_store`` channels each (x = hi + lo) and each of its (tap, chunk)
    pairs becomes three k-blocks: x_hi*w_hi, x_hi*w_lo and x_lo*w_hi (the x_lo*w_lo term is below 2^-22 relative); a plain
    fp16 source gets x*w_hi and x*w_lo.  ``split_out`` (default: = strict) writes the output as (hi, lo) pairs so that a
    strict consumer can read it; a non-strict layer may also produce a split output."""
    if split_out is None:
        split_out = strict
    assert strict or not any(p.split for p in parts), 'a (hi, lo) source needs strict=True'
    if KC is None:
        KC = 64 if all(p.c_store % 64 == 0 for p in parts) else 32
    co_real = parts[0].w.shape[0]
    assert co_store % 16 == 0 and co_store >= co_real
    ws = []
    for p in parts:
        assert p.c_store % KC == 0, (p.c_store, KC)
        w = p.w.detach().to(torch.float32).cpu()
        if w.dim() == 4:
            w = w[:, :, None]
        if out_scale is not None:
            w = w * out_scale.detach().cpu().view(-1, 1, 1, 1, 1)
        ws.append(w)
    osc = _row_scales(ws, co_store)
    kbs, blocks = [], []
    for si, (p, w) in enumerate(zip(parts, ws)):
        co, ci, kd, kh, kw = w.shape
        wp = torch.zeros((co_store, p.c_store, kd, kh, kw), dtype=torch.float32)
        wp[:co, :ci] = w
        if osc is not None:
            wp = wp / osc.view(-1, 1, 1, 1, 1)
        if strict:
            w_hi = wp.to(torch.float16).to(torch.float32)
            w_lo = wp - w_hi
        for q in range(kd):
            for r in range(kh):
                for s in range(kw):
                    for c0 in range(0, p.c_store, KC):
                        blk = wp[:, c0:c0 + KC, q, r, s]
                        if not bool(blk.any()):
                            continue
                        tap = (s * p.dil, r * p.dil, q * p.dil)
                        if not strict:
                            kbs.append(tap + (c0, si)); blocks.append(blk)
                            continue
                        bh, bl = w_hi[:, c0:c0 + KC, q, r, s], w_lo[:, c0:c0 + KC, q, r, s]
                        kbs.append(tap + (c0, si)); blocks.append(bh)                       # x_hi * w_hi
                        if bool(bl.to(torch.float16).any()):
                            kbs.append(tap + (c0, si)); blocks.append(bl)                   # x_hi * w_lo
                        if p.split:
                            kbs.append(tap + (c0 + p.c_store, si)); blocks.append(bh)       # x_lo * w_hi
    if not blocks:      # degenerate all-zero conv: keep one block so the kernel has work
        kbs.append((0, 0, 0, 0, 0)); blocks.append(torch.zeros((co_store, KC)))
    if len(kbs) > _lib.TPZ_TC_MAX_KB:
        raise RuntimeError(f'topaz_b200: conv needs {len(kbs)} k-blocks (max {_lib.TPZ_TC_MAX_KB})')
    wt = torch.stack(blocks).to(torch.float16).contiguous()
    if not bool(torch.isfinite(wt).all()):
        raise RuntimeError('topaz_b200: convolution weights do not fit the fp16 range')
    wt = wt.to(device)
    b = torch.zeros(co_store, dtype=torch.float32)
    if bias is not None:
        b[:co_real] = bias.detach().to(torch.float32).cpu()
    dw = None
    if dot_w is not None:
        dw = torch.zeros(co_store, dtype=torch.float32)
        dw[:co_real] = dot_w.detach().to(torch.float32).cpu().reshape(-1)
        dw = dw.to(device)
    rs = None
    if res_scale is not None:
        rs = torch.ones(co_store, dtype=torch.float32)
        rs[:co_real] = res_scale.detach().to(torch.float32).cpu()
        rs = rs.to(device)
    grids = [(int(p.w.shape[-1]), int(p.w.shape[-2])) for p in parts]
    if lattice is None:
        dils = {p.dil for p, g in zip(parts, grids) if g != (1, 1)}
        lattice = dils.pop() if len(dils) == 1 else (1 if not dils else 0)
    return TcConvPlan(KC=KC, Co=co_store, kblocks=kbs, orgs=[tuple(p.org) for p in parts],
                      c_stores=[(2 if p.split else 1) * p.c_store for p in parts], tapgrids=grids, lattice=lattice,
                      lats=[p.lat for p in parts], phases=[p.phase for p in parts], phase_sel=phase_sel,
                      lat_zs=[p.lat_z for p in parts], lattice_z=lattice_z, phase_z=phase_z,
                      weights=wt, bias=b.to(device),
                      neg_slope=float(neg_slope), dot_w=dw, dot_b=float(dot_b), res_scale=rs,
                      oscale=osc.to(device) if osc is not None else None, split_out=bool(split_out), w_split=bool(strict))


def _static_tc_args(plan: TcConvPlan) -> TpzTcConvArgs:
    """The launch-invariant part of the argument block (k-block table, weights, epilogue constants); built once per
    plan and cached on it - filling the 256-entry table in Python on every launch cost more than the launch itself."""
    a = plan.__dict__.get('_args')
    if a is not None:
        return a
    a = TpzTcConvArgs()
    a.nsrc = len(plan.c_stores)
    for i in range(a.nsrc):
        s = a.src[i]
        s.C = plan.c_stores[i]
        for j in range(3):
            s.org[j] = plan.orgs[i][j]
        s.kw, s.kh = plan.tapgrids[i]
        s.lat = plan.lats[i] if plan.lats else 0
        s.no_phase = 0 if (not plan.phases or plan.phases[i]) else 1
        s.lat_z = plan.lat_zs[i] if plan.lat_zs else 0
    a.weights = plan.weights.data_ptr()
    a.KC = plan.KC
    a.nkb = len(plan.kblocks)
    for i, (dx, dy, dz, c0, si) in enumerate(plan.kblocks):
        k = a.kb[i]
        k.dx, k.dy, k.dz, k.c0, k.src = dx, dy, dz, c0, si
    a.Co = plan.Co
    a.TW, a.TH = plan.TW, plan.TH
    a.lattice = plan.lattice
    a.phase_sel = plan.phase_sel
    a.lattice_z, a.phase_z = plan.lattice_z, plan.phase_z
    a.bias = plan.bias.data_ptr()
    a.neg_slope = plan.neg_slope
    a.oscale = plan.oscale.data_ptr() if plan.oscale is not None else None
    plan.__dict__['_args'] = a
    return a


def fill_tc_args(plan: TcConvPlan, srcs: Sequence[torch.Tensor], out_shape: Tuple[int, int, int, int],
                 out: Optional[torch.Tensor], res: Optional[torch.Tensor] = None,
                 res_org: Tuple[int, int, int] = (0, 0, 0), dot_out: Optional[torch.Tensor] = None,
                 out_coff: int = 0, dot_affine: Optional[torch.Tensor] = None, rng: Optional[torch.Tensor] = None) -> TpzTcConvArgs:
    a = _static_tc_args(plan)
    assert len(srcs) == a.nsrc
    for i, t in enumerate(srcs):
        N, D, H, W, ld = t.shape
        s = a.src[i]
        s.ptr = t.data_ptr(); s.N, s.D, s.H, s.W = N, D, H, W
        s.ld = ld
    a.N, a.Do, a.Ho, a.Wo = out_shape
    if res is not None:
        a.res = res.data_ptr(); a.res_ld = res.shape[4]
        a.res_D, a.res_H, a.res_W = res.shape[1], res.shape[2], res.shape[3]
        for j in range(3):
            a.res_org[j] = res_org[j]
        a.res_scale = plan.res_scale.data_ptr() if plan.res_scale is not None else None
    else:
        a.res = None; a.res_scale = None
    if out is not None:
        a.out = out.data_ptr(); a.out_ld = out.shape[4]; a.out_coff = out_coff
        a.out_lo = plan.Co if plan.split_out else 0
        assert not plan.split_out or (out_coff == 0 and out.shape[4] == 2 * plan.Co)
    else:
        a.out = None; a.out_lo = 0
    a.range = rng.data_ptr() if rng is not None else None
    if dot_out is not None:
        a.dot_w = plan.dot_w.data_ptr(); a.dot_b = plan.dot_b; a.dot_out = dot_out.data_ptr()
        a.dot_affine = dot_affine.data_ptr() if dot_affine is not None else None
    else:
        a.dot_w = None; a.dot_out = None; a.dot_affine = None
    return a


TC_VARIANT = 'auto'      # 'auto' (halo-resident kernel when eligible), 'v1' (per-tap loads), 'v2' (must be eligible)
LAUNCH_COUNT = 0          # kernels launched through this module (bench.py reports it as gpu_launches)
EVENT_HOOK = None         # optional callable(tag) -> (start_event, end_event) recorder used by bench.py


def tc_conv(plan: TcConvPlan, srcs, out_shape, out=None, res=None, res_org=(0, 0, 0), dot_out=None, out_coff=0,
            tag=None, dot_affine=None, rng=None):
    global LAUNCH_COUNT
    a = fill_tc_args(plan, srcs, out_shape, out, res, res_org, dot_out, out_coff, dot_affine, rng)
    fn = {'auto': _lib.lib().tpz_tc_conv, 'v1': _lib.lib().tpz_tc_conv_v1, 'v2': _lib.lib().tpz_tc_conv_v2}[TC_VARIANT]
    hook = EVENT_HOOK(tag) if (EVENT_HOOK is not None and tag is not None) else None
    if hook is not None:
        hook[0].record(torch.cuda.current_stream())
    check(fn(C.byref(a), _stream()))
    if hook is not None:
        hook[1].record(torch.cuda.current_stream())
    LAUNCH_COUNT += 1


def _count(n):
    global LAUNCH_COUNT
    LAUNCH_COUNT += n


# ------------------------------------------------------------------------------------------------
# direct kernels
# ------------------------------------------------------------------------------------------------
_RANGE_WORK = {}


def range_scale(x: torch.Tensor) -> torch.Tensor:
    """Device float[2] = (s, 1/s): power-of-two scale of the fp16 activations chosen from max|x| (tpz_range_scale);
    s = 1 for inputs in the normalised range.  No host synchronisation."""
    key = (x.device.index, torch.cuda.current_stream().cuda_stream)
    work = _RANGE_WORK.get(key)
    if work is None:
        work = _RANGE_WORK[key] = torch.zeros(1, dtype=torch.int32, device=x.device)
    rng = torch.empty(2, dtype=torch.float32, device=x.device)
    _count(2); check(_lib.lib().tpz_range_scale(_ptr(x), x.numel(), _ptr(rng), _ptr(work), _stream()))
    return rng


def conv_first(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], dil: int, pad: int,
               neg_slope: float, out_ld: int, rng: Optional[torch.Tensor] = None, split: bool = False) -> torch.Tensor:
    """x: fp32 [N, D, H, W]; w: fp32 [Co, kd, kh, kw] (device).  Returns fp16 [N, Do, Ho, Wo, out_ld] (``split``: strict
    mode, [.., 2*out_ld] = hi | lo halves)."""
    N, D, H, W = x.shape
    Co, kd, kh, kw = w.shape
    Do = D + 2 * pad - (kd - 1) * dil if kd > 1 else D
    Ho, Wo = H + 2 * pad - (kh - 1) * dil, W + 2 * pad - (kw - 1) * dil
    ld = 2 * out_ld if split else out_ld
    out = torch.empty((N, Do, Ho, Wo, ld), dtype=torch.float16, device=x.device)
    _count(1); check(_lib.lib().tpz_conv_first(_ptr(x), N, D, H, W, _ptr(w), _ptr(bias), Co, kd, kh, kw, dil, pad,
                                    float(neg_slope), 1, _ptr(out), ld, _ptr(rng), out_ld if split else 0, _stream()))
    return out


def im2col3d_first(x: torch.Tensor, k: int, ld: int, rng: Optional[torch.Tensor] = None, split: bool = False) -> torch.Tensor:
    """x: fp32 [N, D, H, W].  Returns fp16 [N, D, H, W, ld] with channel t = tap (dz*k+dy)*k+dx ('same' padding), zero beyond k^3
    (``split``: [.., 2*ld] = hi | lo halves)."""
    N, D, H, W = x.shape
    out = torch.empty((N, D, H, W, 2 * ld if split else ld), dtype=torch.float16, device=x.device)
    _count(1); check(_lib.lib().tpz_im2col3d_first(_ptr(x), N, D, H, W, k, k // 2, _ptr(out), ld, _ptr(rng), ld if split else 0, _stream()))
    return out


def first_tc_supported(k: int, cp: int) -> bool:
    return (k in (3, 5, 7, 11)) and (cp in (32, 64))


def pack_first_tc(w: torch.Tensor, bias: Optional[torch.Tensor], cp: int, device):
    """w: fp32 [Co, k, k] (BN already folded).  Returns (fp16 [KB][cp][64] k-block-major packed taps, fp32 bias [cp])."""
    co, k, _ = w.shape
    kb = (k * k + 63) // 64
    wp = torch.zeros((cp, kb * 64), dtype=torch.float32)
    wp[:co, :k * k] = w.detach().float().cpu().reshape(co, k * k)
    wp = wp.reshape(cp, kb, 64).permute(1, 0, 2).contiguous().to(torch.float16)
    bp = torch.zeros(cp, dtype=torch.float32)
    if bias is not None:
        bp[:co] = bias.detach().float().cpu()
    return wp.to(device), bp.to(device)


def conv_first_tc(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, k: int, pad: int, neg_slope: float,
                  pool: bool = False, rng: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: fp32 [N, H, W] on device.  Returns fp16 [N, 1, Ho, Wo, Cp]: conv k x k (zero padding `pad`) + bias + activation;
    with ``pool`` the 2x2 max-pool that follows is fused and the result is [N, 1, Ho//2, Wo//2, Cp]."""
    N, H, W = x.shape
    cp = w_packed.shape[1]
    Ho, Wo = H + 2 * pad - (k - 1), W + 2 * pad - (k - 1)
    oshape = (N, 1, Ho // 2, Wo // 2, cp) if pool else (N, 1, Ho, Wo, cp)
    out = torch.empty(oshape, dtype=torch.float16, device=x.device)
    if out.numel() == 0:
        return out
    _count(1); check(_lib.lib().tpz_conv_first_tc(_ptr(x), N, H, W, _ptr(w_packed), _ptr(bias), cp, k, pad, float(neg_slope),
                                                  int(pool), _ptr(out), _ptr(rng), _stream()))
    return out


def im2col_first(x: torch.Tensor, k: int, pad: int, ld: int, rng: Optional[torch.Tensor] = None, split: bool = False) -> torch.Tensor:
    """x: fp32 [N, H, W].  Returns fp16 [N, 1, Ho, Wo, ld] with channel t = tap (r*k+s), zero beyond k*k
    (``split``: [.., 2*ld] = hi | lo halves)."""
    N, H, W = x.shape
    Ho, Wo = H + 2 * pad - (k - 1), W + 2 * pad - (k - 1)
    out = torch.empty((N, 1, Ho, Wo, 2 * ld if split else ld), dtype=torch.float16, device=x.device)
    _count(1); check(_lib.lib().tpz_im2col_first(_ptr(x), N, H, W, k, pad, _ptr(out), ld, _ptr(rng), ld if split else 0, _stream()))
    return out


def conv_last(x: torch.Tensor, c_real: int, w: torch.Tensor, bias: float, kdhw, dil: int, pad: int,
              stats: Optional[torch.Tensor] = None, out_scale: float = 1.0, out_shift: float = 0.0,
              rng: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: fp16 [N,D,H,W,ld]; w: fp32 [taps, C] (device) with C = c_real rounded up to 8. Returns fp32 [N,D,H,W]."""
    N, D, H, W, ld = x.shape
    kd, kh, kw = kdhw
    out = torch.empty((N, D, H, W), dtype=torch.float32, device=x.device)
    _count(1); check(_lib.lib().tpz_conv_last(_ptr(x), N, D, H, W, w.shape[1], ld, _ptr(w), float(bias), kd, kh, kw, dil, pad,
                                   float(out_scale), float(out_shift), _ptr(stats), _ptr(out), _ptr(rng), _stream()))
    return out


def conv_generic(x0, c0, x1, c1, w, bias, stride, dil, pad, neg_slope, out_ld, res=None, res_org=0):
    """Validation conv: x0/x1 fp16 NDHWC (x1 optional), w fp32 [Co, C0+C1, kd, kh, kw] device."""
    N, D, H, W, ld0 = x0.shape
    Co, Ci, kd, kh, kw = w.shape
    assert Ci == c0 + c1
    def osz(n, k, three):
        return (n + 2 * pad - (k - 1) * dil - 1) // stride + 1 if three else n
    Do = osz(D, kd, kd > 1); Ho = osz(H, kh, True); Wo = osz(W, kw, True)
    out = torch.zeros((N, Do, Ho, Wo, out_ld), dtype=torch.float16, device=x0.device)
    _count(1); check(_lib.lib().tpz_conv_generic(_ptr(x0), c0, ld0, _ptr(x1), c1, x1.shape[4] if x1 is not None else 0,
                                      N, D, H, W, _ptr(w), _ptr(bias), Co, kd, kh, kw, stride, dil, pad,
                                      float(neg_slope), _ptr(res), res.shape[4] if res is not None else 0, res_org,
                                      _ptr(out), out_ld, Do, Ho, Wo, _stream()))
    return out


def maxpool2(x: torch.Tensor, dims: int, split: bool = False) -> torch.Tensor:
    """2x max-pool (floor); ``split``: channels are [hi | lo] halves (strict mode), the maximum is over hi + lo."""
    N, D, H, W, ld = x.shape
    Do = D // 2 if dims == 3 else D
    out = torch.empty((N, Do, H // 2, W // 2, ld), dtype=torch.float16, device=x.device)
    C_ = ld // 2 if split else ld
    _count(1); check(_lib.lib().tpz_maxpool2(_ptr(x), N, D, H, W, C_, ld, dims, _ptr(out), ld, C_ if split else 0, _stream()))
    return out


def upsample_nearest(x: torch.Tensor, size: Tuple[int, int, int]) -> torch.Tensor:
    N, D, H, W, ld = x.shape
    Do, Ho, Wo = size
    out = torch.empty((N, Do, Ho, Wo, ld), dtype=torch.float16, device=x.device)
    _count(1); check(_lib.lib().tpz_upsample_nearest(_ptr(x), N, D, H, W, ld, ld, Do, Ho, Wo, _ptr(out), ld, 0, _stream()))
    return out


def meanstd(x: torch.Tensor, unbiased: bool) -> torch.Tensor:
    """Device float[2] = (mean, std) of a contiguous fp32 tensor; no host synchronisation."""
    stats = torch.empty(2, dtype=torch.float32, device=x.device)
    work = torch.empty(4, dtype=torch.float64, device=x.device)
    _count(5); check(_lib.lib().tpz_meanstd(_ptr(x), x.numel(), int(unbiased), _ptr(stats), _ptr(work), _stream()))
    return stats


def affine(x: torch.Tensor, stats: torch.Tensor, inverse: bool = False, out: Optional[torch.Tensor] = None):
    y = torch.empty_like(x) if out is None else out
    _count(1); check(_lib.lib().tpz_affine(_ptr(x), x.numel(), _ptr(stats), int(inverse), _ptr(y), _stream()))
    return y


def gemm_f32(A: torch.Tensor, B: torch.Tensor, C: torch.Tensor):
    """C[M][N] = A[M][K] @ B[K][N], contiguous fp32 device tensors, 3xTF32 tensor-core product (K%16==0, N%32==0)."""
    M, K = A.shape
    K2, N = B.shape
    assert K == K2 and tuple(C.shape) == (M, N) and A.is_contiguous() and B.is_contiguous() and C.is_contiguous()
    _count(1); check(_lib.lib().tpz_gemm_f32(_ptr(A), M, K, _ptr(B), N, _ptr(C), _stream()))
    return C


def gmm_sums(x: torch.Tensor, shift: float, sets8, work: Optional[torch.Tensor] = None):
    """One GMM pass over a flat fp32 device tensor for k <= 12 parameter sets -> numpy float64 [k, 7]
    (see include/topaz_b200.h: tpz_gmm_sums)."""
    import numpy as np
    p = np.ascontiguousarray(sets8, dtype=np.float64).reshape(-1, 8)
    k = p.shape[0]
    buf = work if work is not None else torch.empty(12 * 7, dtype=torch.float64, device=x.device)
    _count(1); check(_lib.lib().tpz_gmm_sums(_ptr(x), x.numel(), float(shift), p.ctypes.data, k, _ptr(buf), _stream()))
    return buf[:k * 7].cpu().numpy().reshape(k, 7)


def select_hist(x: torch.Tensor, level: int, prefixes=()):
    """Radix-select histogram pass (tpz_select_hist) -> numpy int64 [4096] (level 0) or [len(prefixes), 4096 | 256]."""
    import numpy as np
    n_pref = len(prefixes)
    bins = 4096 if level < 2 else 256
    rows = 1 if level == 0 else n_pref
    hist = torch.empty(rows * bins, dtype=torch.int32, device=x.device)
    pref = torch.tensor(list(prefixes), dtype=torch.int64).to(torch.int32).to(x.device) if n_pref else None
    _count(1); check(_lib.lib().tpz_select_hist(_ptr(x), x.numel(), int(level), _ptr(pref) if n_pref else None, n_pref,
                                                _ptr(hist), _stream()))
    h = hist.cpu().numpy().astype(np.int64)
    return h if level == 0 else h.reshape(rows, bins)


def to_device(t: torch.Tensor) -> torch.Tensor:
    """Host -> current CUDA device (separate hook so the CPU simulation of the kernels can keep tensors on the host)."""
    return t if t.is_cuda else t.cuda()


def filter_f32(x: torch.Tensor, f: torch.Tensor, bias: float = 0.0) -> torch.Tensor:
    """Same-padded 1->1 fp32 convolution of x [N,D,H,W] with filter f [kd,kh,kw] (both on device)."""
    N, D, H, W = x.shape
    kd, kh, kw = f.shape
    y = torch.empty_like(x)
    _count(1); check(_lib.lib().tpz_filter_f32(_ptr(x), N, D, H, W, _ptr(f), kd, kh, kw, float(bias), _ptr(y), _stream()))
    return y
