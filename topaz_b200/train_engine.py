"""Strided (unfilled) classifier forward + backward on the fp32 training kernels (csrc/tpz_train.cu).

Activations are NHWC fp32; parameters stay in the reference's OIHW layout inside ONE flat buffer (so the
fused Adam step and the multi-GPU gradient all-reduce each touch a single tensor); gradients are written
straight into the matching flat gradient buffer that ``p.grad`` aliases.

Reference call sites replaced: ``score = self.model(X).view(-1)`` (methods.py:103), ``loss.backward()`` (:146).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.nn as nn

import os

from . import _lib, ops
from ._lib import check
from .engine import _feature_blocks, is_filled


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class FlatParams:
    """Flat fp32 parameter / gradient / Adam-moment buffers with per-parameter views."""

    def __init__(self, model: nn.Module):
        self.params = [p for p in model.parameters()]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.offsets = []
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p)
            p.grad = self.flat_g[off:off + k].view_as(p)
            self.offsets.append(off)
            off += k
        self.n = n
        self.step = 0
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)      # device mirror of `step` (graph-replayed Adam)
        self.ptrs = tuple(p.data_ptr() for p in self.params)
        # packed copies of the conv weights for the tensor-core training kernels (refreshed after every update)
        self.packed_off = {}
        descs, poff, mx = [], 0, 0
        for p, off in zip(self.params, self.offsets):
            if p.dim() == 4 and p.shape[1] % 16 == 0 and p.shape[0] % 16 == 0:
                if USE_TC and p.shape[1] % 32 == 0 and p.shape[0] % 32 == 0:
                    continue                                   # served by the tcgen05 kernels (packed_tc below): no mma.sync copy
                co, ci, kh, kw = p.shape
                k = p.numel()
                self.packed_off[id(p)] = (poff, poff + k)
                descs.append((off, poff, poff + k, co, ci, kh * kw))
                poff += 2 * k
                mx = max(mx, k)
        self.packed = torch.empty(max(poff, 1), dtype=torch.float32, device=dev)
        self.ndesc, self.max_elems = len(descs), mx
        import numpy as _np
        rec = _np.zeros(len(descs), dtype=[('src', '<i8'), ('fwd', '<i8'), ('dg', '<i8'), ('co', '<i4'), ('ci', '<i4'),
                                           ('taps', '<i4'), ('pad', '<i4')])
        for i, d in enumerate(descs):
            rec[i] = (d[0], d[1], d[2], d[3], d[4], d[5], 0)
        self.descs = torch.from_numpy(rec.view(_np.uint8).copy()).to(dev) if descs else None
        # tcgen05 kernels: hi / lo planes in the TMA-friendly [tap][chunk][n][32] layout (4 x numel per eligible weight)
        self.packed_tc_off = {}
        descs_tc, toff, mx_tc = [], 0, 0
        for p, off in zip(self.params, self.offsets):
            if p.dim() == 4 and p.shape[1] % 32 == 0 and p.shape[0] % 32 == 0:
                co, ci, kh, kw = p.shape
                k = p.numel()
                self.packed_tc_off[id(p)] = (toff, toff + 2 * k)
                descs_tc.append((off, toff, toff + 2 * k, co, ci, kh * kw))
                toff += 4 * k
                mx_tc = max(mx_tc, k)
        self.packed_tc = torch.empty(max(toff, 1), dtype=torch.float32, device=dev)
        self.ndesc_tc, self.max_elems_tc = len(descs_tc), mx_tc
        rec2 = _np.zeros(len(descs_tc), dtype=rec.dtype)
        for i, d in enumerate(descs_tc):
            rec2[i] = (d[0], d[1], d[2], d[3], d[4], d[5], 0)
        self.descs_tc = torch.from_numpy(rec2.view(_np.uint8).copy()).to(dev) if descs_tc else None
        self.packed_step = -1
        self.packed_versions = None
        # data parallel: replicas must start from identical parameters and optimizer state (nothing else enforces it);
        # rank 0's values win, as with torch's DistributedDataParallel
        dist = _dp()
        if dist is not None:
            dist.broadcast(self.flat_p, 0)

    def versions(self):
        """in-place version counters of the parameters: load_state_dict / p.data.copy_ / a broadcast bump them"""
        return tuple(p._version for p in self.params)

    def set_step(self, step: int):
        self.step = int(step)
        self.step_dev.fill_(int(step))

    def packed_tc_ptrs(self, w):
        """(forward-layout, dgrad-layout) views of the tcgen05 packed copy (hi | lo planes) of conv weight `w`, or None."""
        o = self.packed_tc_off.get(id(w))
        if o is None:
            return None
        return self.packed_tc[o[0]:], self.packed_tc[o[1]:]

    def packed_ptrs(self, w):
        """(forward-layout, dgrad-layout) views of the packed copy of conv weight `w`, or None."""
        o = self.packed_off.get(id(w))
        if o is None:
            return None
        return self.packed[o[0]:], self.packed[o[1]:]

    def valid_for(self, model) -> bool:
        ps = [p for p in model.parameters()]
        return len(ps) == len(self.params) and all(a is b and a.data_ptr() == q for a, b, q in zip(ps, self.params, self.ptrs))

    def ensure_grads(self):
        """Re-attach p.grad views (a user-side optim.zero_grad(set_to_none=True) detaches them)."""
        for p, off in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat_g[off:].data_ptr():
                p.grad = self.flat_g[off:off + p.numel()].view_as(p)


def flat_params(model) -> FlatParams:
    fp = model.__dict__.get('_tpz_flat')
    if fp is None or not fp.valid_for(model):
        old = fp
        fp = FlatParams(model)
        # Same Parameter objects in new storage: the module went through .cpu()/.cuda() (the reference does that around
        # torch.save after EVERY epoch, training.py:600-603).  torch.optim.Adam keeps its state across that (it is keyed
        # by the Parameter objects), so the moments and the step count move over to the new flat buffers.
        if old is not None and len(old.params) == len(fp.params) and all(a is b for a, b in zip(old.params, fp.params)) \
                and old.n == fp.n:
            fp.flat_m.copy_(old.flat_m)
            fp.flat_v.copy_(old.flat_v)
            fp.set_step(old.step)
        model.__dict__['_tpz_flat'] = fp
    return fp


_CUR = {'fp': None}      # FlatParams of the model whose step is running (packed weights for the mma kernels)
USE_MMA = os.environ.get('TPZ_TRAIN_SIMT') is None
# tcgen05 (kind::tf32, 3-pass) training convolutions for channel counts that are multiples of 32; TPZ_TRAIN_TC=0 -> mma.sync
USE_TC = USE_MMA and os.environ.get('TPZ_TRAIN_TC', '1') == '1'


def _repack(fp, force: bool = False):
    """Refresh the packed weight copies after a parameter update (one launch for all layers).  Stale when the optimizer
    stepped OR a parameter was changed in place between steps (load_state_dict, p.data.copy_, a parameter broadcast):
    the in-place version counters catch the latter."""
    vers = fp.versions()
    if (fp.ndesc or fp.ndesc_tc) and (force or fp.packed_step != fp.step or fp.packed_versions != vers):
        if fp.ndesc:
            ops._count(1)
            check(_lib.lib().tpz_train_repack(_p(fp.flat_p), _p(fp.descs), fp.ndesc, fp.max_elems, _p(fp.packed), _s()))
        if USE_TC and fp.ndesc_tc:
            ops._count(1)
            check(_lib.lib().tpz_train_repack_tc(_p(fp.flat_p), _p(fp.descs_tc), fp.ndesc_tc, fp.max_elems_tc, _p(fp.packed_tc), _s()))
        fp.packed_step = fp.step
        fp.packed_versions = vers


def _packed(w):
    fp = _CUR['fp']
    if not USE_MMA or fp is None:
        return None
    return fp.packed_ptrs(w)


def _packed_tc(w):
    fp = _CUR['fp']
    if not USE_TC or fp is None:
        return None
    return fp.packed_tc_ptrs(w)


def _conv_fwd(x, w, b, stride, dil, org, Ho, Wo, relu, res=None, res_org=0, res_stride=1):
    N, H, W, Ci = x.shape
    Co, _, kh, kw = w.shape
    y = torch.empty((N, Ho, Wo, Co), dtype=torch.float32, device=x.device)
    ops._count(1)
    if USE_MMA and Ci == 1 and Co in (32, 64) and org == 0 and dil == 1 and res is None and kh == kw:
        # 7 x 7 / 32 channels: im2col tile built in shared memory + tcgen05 (returns -1 for other shapes -> CUDA-core kernel)
        rc = _lib.lib().tpz_first_fwd_tc(_p(x), N, H, W, _p(w), _p(b), Co, kh, stride, int(relu), _p(y), Ho, Wo, _s()) if USE_TC else -1
        if rc == -1:
            rc = _lib.lib().tpz_first_fwd_f32(_p(x), N, H, W, _p(w), _p(b), Co, kh, stride, int(relu), _p(y), Ho, Wo, _s())
        check(rc)
        return y
    if USE_MMA and Co == 1 and kh == 1 and kw == 1 and stride == 1 and org == 0 and res is None and not relu and Ci % 4 == 0:
        check(_lib.lib().tpz_cls_fwd_f32(_p(x), N * H * W, Ci, _p(w), _p(b), _p(y), _s()))       # classifier head: one warp per pixel
        return y
    pt = _packed_tc(w)
    if pt is not None and Ci % 32 == 0 and Co % 32 == 0:
        check(_lib.lib().tpz_conv_fwd_tc(_p(x), N, H, W, Ci, _p(pt[0]), _p(b), Co, kh, kw, stride, dil, org, _p(res),
                                         res.shape[1] if res is not None else 0, res.shape[2] if res is not None else 0,
                                         res_org, res_stride, int(relu), _p(y), Ho, Wo, _s()))
        return y
    pk = _packed(w)
    if pk is not None and Ci % 16 == 0 and Co % 32 == 0:
        check(_lib.lib().tpz_conv_fwd_mma(_p(x), N, H, W, Ci, _p(pk[0]), _p(b), Co, kh, kw, stride, dil, org, _p(res),
                                          res.shape[1] if res is not None else 0, res.shape[2] if res is not None else 0,
                                          res_org, res_stride, int(relu), _p(y), Ho, Wo, _s()))
        return y
    check(_lib.lib().tpz_conv_fwd_f32(_p(x), N, H, W, Ci, _p(w), _p(b), Co, kh, kw, stride, dil, org, _p(res),
                                      res.shape[1] if res is not None else 0, res.shape[2] if res is not None else 0,
                                      res_org, res_stride, int(relu), _p(y), Ho, Wo, _s()))
    return y


def _conv_dgrad(dy, w, stride, dil, org, H, W, mask=None, accumulate=False, out=None, res=None, res_org=0):
    """dx = mask(dgrad(dy) [+ out] [+ res embedded at (res_org, res_org)]); `res` is the gradient arriving through a cropped
    identity skip (fused into the halo-resident kernel's epilogue where that kernel applies)."""
    N, Ho, Wo, Co = dy.shape
    _, Ci, kh, kw = w.shape
    dx = out if out is not None else torch.empty((N, H, W, Ci), dtype=torch.float32, device=dy.device)
    ops._count(1)
    pt = _packed_tc(w)
    if pt is not None and Co % 32 == 0 and Ci % 32 == 0:
        if res is not None:
            check(_lib.lib().tpz_conv_dgrad_tc_res(_p(dy), N, Ho, Wo, Co, _p(pt[1]), Ci, kh, kw, stride, dil, org, _p(mask),
                                                   int(accumulate), _p(res), res.shape[1], res.shape[2], res_org, _p(dx), H, W, _s()))
            return dx
        check(_lib.lib().tpz_conv_dgrad_tc(_p(dy), N, Ho, Wo, Co, _p(pt[1]), Ci, kh, kw, stride, dil, org, _p(mask),
                                           int(accumulate), _p(dx), H, W, _s()))
        return dx
    if res is not None:                                   # no fused form on the mma.sync / fp32 paths: three passes
        dx = _conv_dgrad(dy, w, stride, dil, org, H, W, mask=None, accumulate=accumulate, out=out)
        _crop_add(dx, res, res_org, 1)
        if mask is not None:
            _relu_bwd(dx, mask)
        return dx
    pk = _packed(w)
    if pk is not None and Co % 16 == 0 and Ci % 32 == 0:
        check(_lib.lib().tpz_conv_dgrad_mma(_p(dy), N, Ho, Wo, Co, _p(pk[1]), Ci, kh, kw, stride, dil, org, _p(mask),
                                            int(accumulate), _p(dx), H, W, _s()))
        return dx
    check(_lib.lib().tpz_conv_dgrad_f32(_p(dy), N, Ho, Wo, Co, _p(w), Ci, kh, kw, stride, dil, org, _p(mask),
                                        int(accumulate), _p(dx), H, W, _s()))
    return dx


def _conv_wgrad(x, dy, w_grad, b_grad, stride, dil, org):
    N, H, W, Ci = x.shape
    _, Ho, Wo, Co = dy.shape
    kh, kw = w_grad.shape[2], w_grad.shape[3]
    ops._count(2 if b_grad is not None else 1)
    if USE_MMA and Ci == 1 and Co in (32, 64) and org == 0 and dil == 1 and kh == kw and kh * kw <= (256 // Co) * 16:
        rc = _lib.lib().tpz_first_wgrad_tc(_p(x), N, H, W, _p(dy), Ho, Wo, Co, kh, stride, _p(w_grad), _p(b_grad), _s()) if USE_TC else -1
        if rc == 0:
            return                                             # the bias gradient rode along
        if rc == -1:
            rc = _lib.lib().tpz_first_wgrad_f32(_p(x), N, H, W, _p(dy), Ho, Wo, Co, kh, stride, _p(w_grad), _s())
        check(rc)
        if b_grad is not None:
            check(_lib.lib().tpz_bias_grad_f32(_p(dy), N * Ho * Wo, Co, _p(b_grad), _s()))
        return
    if USE_TC and Ci % 32 == 0 and Co % 32 == 0:
        # the bias gradient rides along in the halo-resident wgrad kernel (the library runs the separate reducer otherwise)
        check(_lib.lib().tpz_conv_wgrad_tc_bias(_p(x), N, H, W, Ci, _p(dy), Ho, Wo, Co, kh, kw, stride, dil, org, _p(w_grad),
                                                _p(b_grad), _s()))
        return
    if USE_MMA and Ci % 16 == 0 and Co % 16 == 0:
        check(_lib.lib().tpz_conv_wgrad_mma(_p(x), N, H, W, Ci, _p(dy), Ho, Wo, Co, kh, kw, stride, dil, org, _p(w_grad), _s()))
        if b_grad is not None:      # bias gradient: the fp32 kernel with a zero-tap weight pass is not needed; reuse its reducer
            check(_lib.lib().tpz_bias_grad_f32(_p(dy), N * Ho * Wo, Co, _p(b_grad), _s()))
        return
    check(_lib.lib().tpz_conv_wgrad_f32(_p(x), N, H, W, Ci, _p(dy), Ho, Wo, Co, kh, kw, stride, dil, org, _p(w_grad),
                                        _p(b_grad), _s()))


def _cls_bwd(x, g, w, w_grad, b_grad, masked: bool):
    """backward of the classifier head (1x1 conv C -> 1): returns dx, accumulates w_grad / b_grad; `masked` fuses the ReLU mask of
    the layer that produced x."""
    N, H, W, Ci = x.shape
    if USE_MMA and Ci % 4 == 0 and Ci <= 8192:
        dx = torch.empty_like(x)
        ops._count(1)
        check(_lib.lib().tpz_cls_bwd_f32(_p(x), N * H * W, Ci, _p(w), _p(g), int(masked), _p(dx), _p(w_grad), _p(b_grad), _s()))
        return dx
    _conv_wgrad(x, g, w_grad, b_grad, 1, 1, 0)
    return _conv_dgrad(g, w, 1, 1, 0, H, W, mask=x if masked else None)


def _relu_bwd(dy, y):
    ops._count(1)
    check(_lib.lib().tpz_relu_bwd_f32(_p(dy), _p(y), dy.numel(), _s()))


def _crop_add(dx, g, org, stride):
    N, H, W, Cc = dx.shape
    ops._count(1)
    check(_lib.lib().tpz_crop_add_f32(_p(dx), N, H, W, Cc, _p(g), g.shape[1], g.shape[2], org, stride, _s()))


class _BnWorkspace:
    """fp64 per-channel sum slots of one training step (forward and backward statistics of every BatchNorm layer),
    zeroed by ONE fill launch; slices are handed out in call order."""

    def __init__(self, n_doubles: int, device):
        ops._count(1)
        self.buf = torch.zeros(max(n_doubles, 1), dtype=torch.float64, device=device)
        self.used = 0

    def take(self, n: int) -> torch.Tensor:
        if self.used + n > self.buf.numel():          # e.g. a second forward on the same tape: fall back to a fresh slot
            ops._count(1)
            return torch.zeros(n, dtype=torch.float64, device=self.buf.device)
        out = self.buf[self.used:self.used + n]
        self.used += n
        return out


def _dp():
    """torch.distributed when this process is one rank of a data-parallel job (see methods._run), else None."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def _bn_stats(x, sums):
    ops._count(1)
    check(_lib.lib().tpz_bn_stats_f32(_p(x), x.numel() // x.shape[-1], x.shape[-1], _p(sums), _s()))


def _bn_fwd(x, sums, count, gamma, beta, eps, momentum, running_mean, running_var, relu, save):
    y = torch.empty_like(x)
    ops._count(1)
    check(_lib.lib().tpz_bn_fwd_f32(_p(x), x.numel() // x.shape[-1], x.shape[-1], _p(sums), int(count), _p(gamma), _p(beta),
                                    float(eps), float(momentum), _p(running_mean), _p(running_var), int(relu), _p(y),
                                    _p(save), _s()))
    return y


def _bn_bwd_reduce(g, x, save, sums):
    ops._count(1)
    check(_lib.lib().tpz_bn_bwd_reduce_f32(_p(g), _p(x), x.numel() // x.shape[-1], x.shape[-1], _p(save), _p(sums), _s()))


def _bn_bwd(g, x, save, sums, count, gamma, local_sums, dgamma, dbeta):
    """in place: g <- d(loss)/d(bn input)"""
    ops._count(1)
    check(_lib.lib().tpz_bn_bwd_f32(_p(g), _p(x), x.numel() // x.shape[-1], x.shape[-1], _p(save), _p(sums), int(count),
                                    _p(gamma), _p(local_sums), _p(dgamma), _p(dbeta), _p(g), _s()))


def _bn_forward(c, bn, relu, ws):
    """nn.BatchNorm2d.forward (+ReLU) on the NHWC conv output `c` (reference resnet.py:101-104,186-187,201-203).
    Training mode: statistics of the minibatch -- of the GLOBAL minibatch under data parallelism (the raw fp64 sums are
    all-reduced) -- and the running-statistics update; eval mode: the running statistics.  Returns (y, save, count)."""
    C_ = c.shape[-1]
    save = torch.empty(2 * C_, dtype=torch.float32, device=c.device)
    if bn.training or bn.running_mean is None:
        P = c.numel() // C_
        sums = ws.take(2 * C_)
        _bn_stats(c, sums)
        dist = _dp()
        count = P
        if dist is not None:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
            count = P * dist.get_world_size()
        if count <= 1:
            raise ValueError('Expected more than 1 value per channel when training, got input size ' + str(list(c.shape)))
        track = bn.track_running_stats and bn.running_mean is not None
        if track and bn.momentum is None:
            raise NotImplementedError('topaz_b200: BatchNorm momentum=None (cumulative average) is not supported')
        y = _bn_fwd(c, sums, count, bn.weight, bn.bias, bn.eps, bn.momentum if track else 0.0,
                    bn.running_mean if track else None, bn.running_var if track else None, relu, save)
        if track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        return y, save, count
    save[:C_] = bn.running_mean
    save[C_:] = torch.rsqrt(bn.running_var.float() + bn.eps)
    y = _bn_fwd(c, None, 0, bn.weight, bn.bias, bn.eps, 0.0, None, None, relu, save)
    return y, save, 0


def _bn_backward(g, c, save, count, bn, ws):
    """Gradient through training-mode BatchNorm: in place g <- d/d(c); accumulates bn.weight.grad / bn.bias.grad."""
    if count == 0:
        raise RuntimeError('topaz_b200: backward through an eval-mode BatchNorm layer')
    C_ = c.shape[-1]
    local = ws.take(2 * C_)
    _bn_bwd_reduce(g, c, save, local)
    sums = local
    dist = _dp()
    if dist is not None:
        sums = local.clone()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    _bn_bwd(g, c, save, sums, count, bn.weight, local, bn.weight.grad if bn.weight is not None else None,
            bn.bias.grad if bn.bias is not None else None)
    return g


def _osz(n, k, dil, stride):
    return (n - (k - 1) * dil - 1) // stride + 1


_DROPOUT = {'calls': 0}      # Philox offset: every dropout launch of the process draws from its own stream position


def _dropout_fwd(x, p):
    """nn.Dropout(p) in training: returns (y, keep-mask uint8).  Seeded by torch's default generator (torch.manual_seed)."""
    y = torch.empty_like(x)
    mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    _DROPOUT['calls'] += 1
    dist = _dp()
    rank = dist.get_rank() if dist is not None else 0          # data-parallel ranks draw independent masks for their shards
    seed = (torch.initial_seed() + rank * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    ops._count(1)
    check(_lib.lib().tpz_dropout_fwd_f32(_p(x), x.numel(), float(p), seed, _DROPOUT['calls'] * 4, _p(y), _p(mask), _s()))
    return y, mask


def _dropout_bwd(g, mask, p):
    """in place: g <- g * mask / (1 - p)"""
    ops._count(1)
    check(_lib.lib().tpz_dropout_bwd_f32(_p(g), _p(mask), g.numel(), float(p), _s()))
    return g


def _is_relu(act) -> bool:
    return isinstance(act, nn.ReLU)


def _check_trainable(blocks):
    """ReLU everywhere (the ResNets), or -- in plain conv blocks (conv31/63/127, basic.py:16) -- PReLU with ONE learnable
    slope / LeakyReLU."""
    for b in blocks:
        if b['kind'] == 'dropout':
            continue
        if b['kind'] == 'conv':
            act = b['act']
            if isinstance(act, nn.PReLU) and act.weight.numel() != 1:
                raise NotImplementedError('topaz_b200: per-channel PReLU is not supported')
            if not isinstance(act, (nn.ReLU, nn.PReLU, nn.LeakyReLU)):
                raise NotImplementedError(f'topaz_b200: activation {type(act).__name__} is not supported by the B200 training path')
        elif not (_is_relu(b['act0']) and _is_relu(b['act1'])):
            raise NotImplementedError('topaz_b200: residual blocks train with ReLU only on the B200 path')


def _slope_args(act):
    """(device pointer of the learnable slope | None, constant slope) for tpz_act_*"""
    if isinstance(act, nn.PReLU):
        return act.weight, 0.0
    return None, float(act.negative_slope)


def _act_fwd(v, act):
    """y = PReLU / LeakyReLU(v) (new tensor; v is kept for the backward)."""
    y = torch.empty_like(v)
    sp, sc = _slope_args(act)
    ops._count(1)
    check(_lib.lib().tpz_act_fwd_f32(_p(v), v.numel(), _p(sp), sc, _p(y), _s()))
    return y


def _act_bwd(g, v, act):
    """in place: g <- d/d(v); accumulates the slope gradient of a PReLU."""
    sp, sc = _slope_args(act)
    ops._count(1)
    check(_lib.lib().tpz_act_bwd_f32(_p(g), _p(v), v.numel(), _p(sp), sc, _p(sp.grad) if sp is not None else None, _s()))
    return g


def _forward(model_features, classifier, x: torch.Tensor, save: bool):
    """x: [B,H,W,1] fp32.  Returns (score [B] or features NHWC, tape)."""
    blocks = _feature_blocks(model_features, slopes=False, dropout=True)
    _check_trainable(blocks)
    tape = []
    cur = x
    n_bn = sum(2 * 2 * bn.num_features for blk in blocks for bn in (blk.get('bn'), blk.get('bn0'), blk.get('bn1'))
               if bn is not None)
    ws = _BnWorkspace(n_bn, x.device) if n_bn else None
    for blk in blocks:
        N, H, W, _ = cur.shape
        if blk['kind'] == 'dropout':
            y, keep = _dropout_fwd(cur, blk['p'])
            # 'relu' is inherited: the consumer's fused mask (input > 0) then covers ReLU AND dropped elements at once
            tape.append(dict(kind='dropout', p=blk['p'], mask=keep, relu=tape[-1].get('relu', True) if tape else True))
            cur = y
        elif blk['kind'] == 'conv':
            w, b, bn = blk['w'], blk['b'], blk.get('bn')
            k = w.shape[-1]
            Ho, Wo = _osz(H, k, blk['dil'], blk['stride']), _osz(W, k, blk['dil'], blk['stride'])
            relu = _is_relu(blk['act'])
            y = _conv_fwd(cur, w, b, blk['stride'], blk['dil'], 0, Ho, Wo, relu=relu and bn is None)
            rec = dict(kind='conv', x=cur, y=y, w=w, b=b, stride=blk['stride'], dil=blk['dil'], relu=relu)
            if bn is not None:
                z, sv, cnt = _bn_forward(y, bn, relu, ws)
                rec.update(bn=bn, c=y, save=sv, count=cnt, y=z)
                y = z
            if not relu:                              # PReLU / LeakyReLU: separate pass, pre-activation kept for the backward
                z = _act_fwd(y, blk['act'])
                rec.update(act=blk['act'], v=y, y=z)
                y = z
            tape.append(rec)
            cur = y
        else:
            w0, b0, w1, b1 = blk['w0'], blk['b0'], blk['w1'], blk['b1']
            bn0, bn1 = blk.get('bn0'), blk.get('bn1')
            d0, d1, s = blk['d0'], blk['d1'], blk['stride']
            H1, W1 = _osz(H, 3, d0, 1), _osz(W, 3, d0, 1)
            h = _conv_fwd(cur, w0, b0, 1, d0, 0, H1, W1, relu=bn0 is None)
            rec = dict(kind='resid', x=cur, w0=w0, b0=b0, w1=w1, b1=b1, proj=blk['proj'], d0=d0, d1=d1, stride=s)
            if bn0 is not None:
                hc = h
                h, sv, cnt = _bn_forward(hc, bn0, True, ws)
                rec.update(bn0=bn0, c0=hc, save0=sv, count0=cnt)
            Ho, Wo = _osz(H1, 3, d1, s), _osz(W1, 3, d1, s)
            edge = d0 + d1
            pr = None
            if blk['proj'] is not None:
                pr = _conv_fwd(cur, blk['proj'], None, s, 1, edge, Ho, Wo, relu=False)
                y = _conv_fwd(h, w1, b1, s, d1, 0, Ho, Wo, relu=bn1 is None, res=pr, res_org=0, res_stride=1)
            else:
                y = _conv_fwd(h, w1, b1, s, d1, 0, Ho, Wo, relu=bn1 is None, res=cur, res_org=edge, res_stride=s)
            if bn1 is not None:
                yc = y
                y, sv, cnt = _bn_forward(yc, bn1, True, ws)
                rec.update(bn1=bn1, c1=yc, save1=sv, count1=cnt)
            rec.update(h=h, y=y, edge=edge)
            tape.append(rec)
            cur = y
    if ws is not None and tape:
        tape[0]['bn_ws'] = ws
    if classifier is None:
        return cur, tape
    N, H, W, _ = cur.shape
    sc = _conv_fwd(cur, classifier.weight, classifier.bias, 1, 1, 0, H, W, relu=False)
    tape.append(dict(kind='cls', x=cur, y=sc, w=classifier.weight, b=classifier.bias))
    return sc, tape


def _prep_input(x: torch.Tensor) -> torch.Tensor:
    ops.require_cuda(x, 'classifier input')
    if x.dim() == 4:
        if x.shape[1] != 1:
            raise ValueError('topaz_b200: expected a single input channel')
        x = x[:, 0]
    return x.contiguous().float().unsqueeze(-1)      # [B,H,W,1] (NHWC with C=1 shares memory with [B,H,W])


def classifier_forward(model, x: torch.Tensor) -> torch.Tensor:
    """Unfilled LinearClassifier.forward: [B,(1,)H,W] -> [B,1,Ho,Wo].  When grad mode is on and the model is in
    train() mode, the activation tape is kept on the model for ``backward``."""
    if is_filled(model.features):
        raise RuntimeError('train_engine.classifier_forward called on a filled model')
    xi = _prep_input(x)
    save = torch.is_grad_enabled() and model.training
    fp = None
    if save:
        fp = flat_params(model)         # make sure params / grads live in the flat buffers before taping pointers
        _repack(fp, force=torch.cuda.is_current_stream_capturing() if x.is_cuda else False)
    _CUR['fp'] = fp
    with torch.no_grad():
        sc, tape = _forward(model.features, model.classifier, xi, save)
    model.__dict__['_tpz_tape'] = tape if save else None
    return sc.permute(0, 3, 1, 2).contiguous()


def features_forward(features, x: torch.Tensor) -> torch.Tensor:
    xi = _prep_input(x)
    with torch.no_grad():
        z, _ = _forward(features, None, xi, False)
    return z.permute(0, 3, 1, 2).contiguous()


def backward(model, dscore: torch.Tensor, on_suffix_done=None):
    """Back-propagate d(loss)/d(score) ([B] fp32, device) through the taped forward; accumulates into p.grad
    (the flat gradient buffer).  ``on_suffix_done(offset)`` is called after each block's backward with the flat-buffer
    offset from which every gradient is final (the buffer is laid out in forward order, the backward runs in reverse):
    data-parallel training starts the all-reduce of that suffix while earlier layers are still back-propagating."""
    tape = model.__dict__.get('_tpz_tape')
    if not tape:
        raise RuntimeError('topaz_b200: backward() without a taped forward (call model(X) in train() mode first)')
    fp = flat_params(model)
    fp.ensure_grads()
    _CUR['fp'] = fp
    g = None
    ws = tape[0].get('bn_ws')
    with torch.no_grad():
        for idx in range(len(tape) - 1, -1, -1):
            rec = tape[idx]
            # the producer of this op's input applies ReLU: its mask (input > 0) is fused into the data-gradient kernel;
            # a PReLU / LeakyReLU producer gets the unmasked gradient and runs its own activation backward
            in_relu = idx > 0 and tape[idx - 1].get('relu', True)
            if rec['kind'] == 'cls':
                x = rec['x']
                N, H, W, _ = x.shape
                g = dscore.contiguous().view(N, H, W, 1)
                g = _cls_bwd(x, g, rec['w'], rec['w'].grad, rec['b'].grad if rec['b'] is not None else None, in_relu)
            elif rec['kind'] == 'dropout':
                g = _dropout_bwd(g, rec['mask'], rec['p'])
            elif rec['kind'] == 'conv':
                x = rec['x']
                if rec.get('act') is not None:                                # g: d/d(activation output) -> d/d(pre-activation)
                    g = _act_bwd(g, rec['v'], rec['act'])
                if rec.get('bn') is not None:                                 # g: d/d(bn output) -> d/d(conv output)
                    g = _bn_backward(g, rec['c'], rec['save'], rec['count'], rec['bn'], ws)
                _conv_wgrad(x, g, rec['w'].grad, rec['b'].grad if rec['b'] is not None else None, rec['stride'], rec['dil'], 0)
                if x.shape[3] == 1 and rec is tape[0]:
                    g = None                                                  # network input: no data gradient needed
                else:
                    g = _conv_dgrad(g, rec['w'], rec['stride'], rec['dil'], 0, x.shape[1], x.shape[2], mask=x if in_relu else None)
            else:
                x, h = rec['x'], rec['h']
                s, d0, d1, edge = rec['stride'], rec['d0'], rec['d1'], rec['edge']
                if rec.get('bn1') is not None:                                # bn1 sits after the skip addition (resnet.py:200-203)
                    g = _bn_backward(g, rec['c1'], rec['save1'], rec['count1'], rec['bn1'], ws)
                _conv_wgrad(h, g, rec['w1'].grad, rec['b1'].grad if rec['b1'] is not None else None, s, d1, 0)
                dh = _conv_dgrad(g, rec['w1'], s, d1, 0, h.shape[1], h.shape[2], mask=h)
                if rec.get('bn0') is not None:
                    dh = _bn_backward(dh, rec['c0'], rec['save0'], rec['count0'], rec['bn0'], ws)
                _conv_wgrad(x, dh, rec['w0'].grad, rec['b0'].grad if rec['b0'] is not None else None, 1, d0, 0)
                # dx = relu'(x) * (dgrad(conv0) + gradient through the skip); x is the previous layer's ReLU output
                if rec['proj'] is not None:
                    _conv_wgrad(x, g, rec['proj'].grad, None, s, 1, edge)
                    dx = _conv_dgrad(g, rec['proj'], s, 1, edge, x.shape[1], x.shape[2])
                    dx = _conv_dgrad(dh, rec['w0'], 1, d0, 0, x.shape[1], x.shape[2], mask=x, accumulate=True, out=dx)
                elif s == 1:
                    dx = _conv_dgrad(dh, rec['w0'], 1, d0, 0, x.shape[1], x.shape[2], mask=x, res=g, res_org=edge)
                else:
                    dx = _conv_dgrad(dh, rec['w0'], 1, d0, 0, x.shape[1], x.shape[2])
                    _crop_add(dx, g, edge, s)
                    _relu_bwd(dx, x)
                g = dx
            if on_suffix_done is not None:
                on_suffix_done(_block_offset(fp, rec))
    model.__dict__['_tpz_tape'] = None
    return fp


def _block_offset(fp, rec) -> int:
    """smallest flat-buffer offset among the parameters of a tape record (incl. its BatchNorm / PReLU parameters)"""
    ptrs = set()
    for k in ('w', 'b', 'w0', 'b0', 'w1', 'b1', 'proj'):
        t = rec.get(k)
        if t is not None:
            ptrs.add(t.data_ptr())
    for k in ('bn', 'bn0', 'bn1'):
        m = rec.get(k)
        if m is not None:
            ptrs.update(p.data_ptr() for p in m.parameters())
    act = rec.get('act')
    if isinstance(act, nn.PReLU):
        ptrs.add(act.weight.data_ptr())
    offs = [off for p, off in zip(fp.params, fp.offsets) if p.data_ptr() in ptrs]
    return min(offs) if offs else fp.n


def ge_loss_grad(scores, labels, pi, slack, lo, hi, dscore, out5):
    """Fused GE-binomial loss / metrics / d(loss)/d(score) (reference methods.py:103-151)."""
    ops._count(1)
    check(_lib.lib().tpz_ge_binomial_loss_grad(_p(scores), _p(labels), scores.numel(), float(pi), float(slack), lo, hi,
                                               _p(dscore), _p(out5), _s()))


def pu_objective_loss_grad(scores, labels, mode, pi, slack, momentum, aux_in, lo, hi, dscore, out6):
    """PN (mode 0) / GE_KL (1) / PU (2) loss, metrics and d(loss)/d(score) (reference methods.py:25-74,168-322)."""
    ops._count(1)
    check(_lib.lib().tpz_pu_objective_loss_grad(_p(scores), _p(labels), scores.numel(), int(mode), float(pi), float(slack),
                                                float(momentum), float(aux_in), lo, hi, _p(dscore), _p(out6), _s()))


def adam_step(fp: FlatParams, lr, b1, b2, eps, l2):
    """Fused Adam on the flat buffers + L2 term + gradient zeroing (reference methods.py:153-160).  The step count lives on
    the device (fp.step_dev, advanced by the launch itself) so that the call can be replayed from a CUDA graph; fp.step is
    its host mirror."""
    fp.step += 1
    ops._count(2)
    check(_lib.lib().tpz_adam_step_dev(_p(fp.flat_p), _p(fp.flat_g), _p(fp.flat_m), _p(fp.flat_v), fp.n, lr, b1, b2, eps,
                                       _p(fp.step_dev), float(l2), 1.0, _s()))


def set_tf32(single_pass: bool) -> bool:
    """Select the training-conv precision: False (default) = 3xTF32 (fp32-level), True = single-pass TF32 (what the
    reference's cuDNN path computes with allow_tf32=True).  Returns the previous setting."""
    return bool(_lib.lib().tpz_train_set_tf32(1 if single_pass else 0))


def read_back(dev_vec: torch.Tensor, host_vec: torch.Tensor):
    """Single host synchronisation of a training step: copy the 5 loss/metric floats to pinned memory."""
    host_vec.copy_(dev_vec, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host_vec.tolist()
