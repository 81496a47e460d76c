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
        self.ptrs = tuple(p.data_ptr() for p in self.params)
        # packed copies of the conv weights for the tensor-core training kernels (refreshed after every update)
        self.packed_off = {}
        descs, poff, mx = [], 0, 0
        for p, off in zip(self.params, self.offsets):
            if p.dim() == 4 and p.shape[1] % 16 == 0 and p.shape[0] % 16 == 0:
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
        self.packed_step = -1

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
        fp = FlatParams(model)
        model.__dict__['_tpz_flat'] = fp
    return fp


_CUR = {'fp': None}      # FlatParams of the model whose step is running (packed weights for the mma kernels)
USE_MMA = os.environ.get('TPZ_TRAIN_SIMT') is None


def _repack(fp):
    """Refresh the packed weight copies after a parameter update (one launch for all layers)."""
    if fp.ndesc and fp.packed_step != fp.step:
        ops._count(1)
        check(_lib.lib().tpz_train_repack(_p(fp.flat_p), _p(fp.descs), fp.ndesc, fp.max_elems, _p(fp.packed), _s()))
        fp.packed_step = fp.step


def _packed(w):
    fp = _CUR['fp']
    if not USE_MMA or fp is None:
        return None
    return fp.packed_ptrs(w)


def _conv_fwd(x, w, b, stride, dil, org, Ho, Wo, relu, res=None, res_org=0, res_stride=1):
    N, H, W, Ci = x.shape
    Co, _, kh, kw = w.shape
    y = torch.empty((N, Ho, Wo, Co), dtype=torch.float32, device=x.device)
    ops._count(1)
    if USE_MMA and Ci == 1 and Co in (32, 64) and org == 0 and dil == 1 and res is None and kh == kw:
        check(_lib.lib().tpz_first_fwd_f32(_p(x), N, H, W, _p(w), _p(b), Co, kh, stride, int(relu), _p(y), Ho, Wo, _s()))
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


def _conv_dgrad(dy, w, stride, dil, org, H, W, mask=None, accumulate=False, out=None):
    N, Ho, Wo, Co = dy.shape
    _, Ci, kh, kw = w.shape
    dx = out if out is not None else torch.empty((N, H, W, Ci), dtype=torch.float32, device=dy.device)
    ops._count(1)
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
        check(_lib.lib().tpz_first_wgrad_f32(_p(x), N, H, W, _p(dy), Ho, Wo, Co, kh, stride, _p(w_grad), _s()))
        if b_grad is not None:
            check(_lib.lib().tpz_bias_grad_f32(_p(dy), N * Ho * Wo, Co, _p(b_grad), _s()))
        return
    if USE_MMA and Ci % 16 == 0 and Co % 16 == 0:
        check(_lib.lib().tpz_conv_wgrad_mma(_p(x), N, H, W, Ci, _p(dy), Ho, Wo, Co, kh, kw, stride, dil, org, _p(w_grad), _s()))
        if b_grad is not None:      # bias gradient: the fp32 kernel with a zero-tap weight pass is not needed; reuse its reducer
            check(_lib.lib().tpz_bias_grad_f32(_p(dy), N * Ho * Wo, Co, _p(b_grad), _s()))
        return
    check(_lib.lib().tpz_conv_wgrad_f32(_p(x), N, H, W, Ci, _p(dy), Ho, Wo, Co, kh, kw, stride, dil, org, _p(w_grad),
                                        _p(b_grad), _s()))


def _relu_bwd(dy, y):
    ops._count(1)
    check(_lib.lib().tpz_relu_bwd_f32(_p(dy), _p(y), dy.numel(), _s()))


def _crop_add(dx, g, org, stride):
    N, H, W, Cc = dx.shape
    ops._count(1)
    check(_lib.lib().tpz_crop_add_f32(_p(dx), N, H, W, Cc, _p(g), g.shape[1], g.shape[2], org, stride, _s()))


def _osz(n, k, dil, stride):
    return (n - (k - 1) * dil - 1) // stride + 1


def _check_trainable(blocks):
    for b in blocks:
        if b.get('bn') is not None or b.get('bn0') is not None or b.get('bn1') is not None:
            raise NotImplementedError('topaz_b200: BatchNorm classifiers are not supported by the B200 training path '
                                      '(the default `topaz train` models are BN-free pretrained ResNets)')
        for k in ('slope', 'slope0', 'slope1'):
            if k in b and b[k] != 0.0:
                raise NotImplementedError('topaz_b200: only ReLU classifiers are supported by the B200 training path')


def _forward(model_features, classifier, x: torch.Tensor, save: bool):
    """x: [B,H,W,1] fp32.  Returns (score [B] or features NHWC, tape)."""
    blocks = _feature_blocks(model_features)
    _check_trainable(blocks)
    tape = []
    cur = x
    for blk in blocks:
        N, H, W, _ = cur.shape
        if blk['kind'] == 'conv':
            w, b = blk['w'], blk['b']
            k = w.shape[-1]
            Ho, Wo = _osz(H, k, blk['dil'], blk['stride']), _osz(W, k, blk['dil'], blk['stride'])
            y = _conv_fwd(cur, w, b, blk['stride'], blk['dil'], 0, Ho, Wo, relu=True)
            tape.append(dict(kind='conv', x=cur, y=y, w=w, b=b, stride=blk['stride'], dil=blk['dil']))
            cur = y
        else:
            w0, b0, w1, b1 = blk['w0'], blk['b0'], blk['w1'], blk['b1']
            d0, d1, s = blk['d0'], blk['d1'], blk['stride']
            H1, W1 = _osz(H, 3, d0, 1), _osz(W, 3, d0, 1)
            h = _conv_fwd(cur, w0, b0, 1, d0, 0, H1, W1, relu=True)
            Ho, Wo = _osz(H1, 3, d1, s), _osz(W1, 3, d1, s)
            edge = d0 + d1
            pr = None
            if blk['proj'] is not None:
                pr = _conv_fwd(cur, blk['proj'], None, s, 1, edge, Ho, Wo, relu=False)
                y = _conv_fwd(h, w1, b1, s, d1, 0, Ho, Wo, relu=True, res=pr, res_org=0, res_stride=1)
            else:
                y = _conv_fwd(h, w1, b1, s, d1, 0, Ho, Wo, relu=True, res=cur, res_org=edge, res_stride=s)
            tape.append(dict(kind='resid', x=cur, h=h, y=y, w0=w0, b0=b0, w1=w1, b1=b1, proj=blk['proj'], d0=d0, d1=d1,
                             stride=s, edge=edge))
            cur = y
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
        _repack(fp)
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


def backward(model, dscore: torch.Tensor):
    """Back-propagate d(loss)/d(score) ([B] fp32, device) through the taped forward; accumulates into p.grad
    (the flat gradient buffer)."""
    tape = model.__dict__.get('_tpz_tape')
    if not tape:
        raise RuntimeError('topaz_b200: backward() without a taped forward (call model(X) in train() mode first)')
    fp = flat_params(model)
    fp.ensure_grads()
    _CUR['fp'] = fp
    g = None
    with torch.no_grad():
        for rec in reversed(tape):
            if rec['kind'] == 'cls':
                x = rec['x']
                N, H, W, _ = x.shape
                g = dscore.contiguous().view(N, H, W, 1)
                _conv_wgrad(x, g, rec['w'].grad, rec['b'].grad, 1, 1, 0)
                g = _conv_dgrad(g, rec['w'], 1, 1, 0, H, W, mask=x)          # masked by relu of the last feature conv
            elif rec['kind'] == 'conv':
                x = rec['x']
                _conv_wgrad(x, g, rec['w'].grad, rec['b'].grad if rec['b'] is not None else None, rec['stride'], rec['dil'], 0)
                if x.shape[3] == 1 and rec is tape[0]:
                    g = None                                                  # network input: no data gradient needed
                else:
                    g = _conv_dgrad(g, rec['w'], rec['stride'], rec['dil'], 0, x.shape[1], x.shape[2], mask=x)
            else:
                x, h = rec['x'], rec['h']
                s, d0, d1, edge = rec['stride'], rec['d0'], rec['d1'], rec['edge']
                _conv_wgrad(h, g, rec['w1'].grad, rec['b1'].grad if rec['b1'] is not None else None, s, d1, 0)
                dh = _conv_dgrad(g, rec['w1'], s, d1, 0, h.shape[1], h.shape[2], mask=h)
                _conv_wgrad(x, dh, rec['w0'].grad, rec['b0'].grad if rec['b0'] is not None else None, 1, d0, 0)
                dx = _conv_dgrad(dh, rec['w0'], 1, d0, 0, x.shape[1], x.shape[2])
                if rec['proj'] is not None:
                    _conv_wgrad(x, g, rec['proj'].grad, None, s, 1, edge)
                    _conv_dgrad(g, rec['proj'], s, 1, edge, x.shape[1], x.shape[2], accumulate=True, out=dx)
                else:
                    _crop_add(dx, g, edge, s)
                _relu_bwd(dx, x)            # x is the previous layer's ReLU output
                g = dx
    model.__dict__['_tpz_tape'] = None
    return fp


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
    """Fused Adam on the flat buffers + L2 term + gradient zeroing (reference methods.py:153-160)."""
    fp.step += 1
    ops._count(1)
    check(_lib.lib().tpz_adam_step(_p(fp.flat_p), _p(fp.flat_g), _p(fp.flat_m), _p(fp.flat_v), fp.n, lr, b1, b2, eps,
                                   fp.step, float(l2), 1.0, _s()))


def set_tf32(single_pass: bool) -> bool:
    """Select the training-conv precision: False (default) = 3xTF32 (fp32-level), True = single-pass TF32 (what the
    reference's cuDNN path computes with allow_tf32=True).  Returns the previous setting."""
    return bool(_lib.lib().tpz_train_set_tf32(1 if single_pass else 0))


def read_back(dev_vec: torch.Tensor, host_vec: torch.Tensor):
    """Single host synchronisation of a training step: copy the 5 loss/metric floats to pinned memory."""
    host_vec.copy_(dev_vec, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host_vec.tolist()
