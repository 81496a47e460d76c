"""CPU *simulation* of the kernel semantics, used only by the not-gpu tests to validate host logic
(weight packing, k-block tables, offsets, channel padding, BN folding, network wiring) without a GPU.
It interprets exactly the arguments the real launchers receive.  Test infrastructure, never shipped."""
import contextlib
import math

import numpy as np
import torch
import torch.nn.functional as F

from topaz_b200 import ops


def _gather(src, org, d, out_shape, c0, kc):
    """src [N,D,H,W,C] fp16 -> fp32 [N,Do,Ho,Wo,kc] window starting at (org+d) with zero fill."""
    N, Do, Ho, Wo = out_shape
    _, D, H, W, _ = src.shape
    ox, oy, oz = org[0] + d[0], org[1] + d[1], org[2] + d[2]
    out = torch.zeros((N, Do, Ho, Wo, kc), dtype=torch.float32)
    z0, z1 = max(0, oz), min(D, oz + Do)
    y0, y1 = max(0, oy), min(H, oy + Ho)
    x0, x1 = max(0, ox), min(W, ox + Wo)
    if z1 > z0 and y1 > y0 and x1 > x0:
        out[:, z0 - oz:z1 - oz, y0 - oy:y1 - oy, x0 - ox:x1 - ox] = src[:, z0:z1, y0:y1, x0:x1, c0:c0 + kc].float()
    return out


def _gather_lattice(src, org, d, qshape, c0, kc, lat, ph, latz=1, phz=0):
    """A operand of one k-block on an output lattice: lattice point (qy,qx) reads source pixel q*lat + ph + org + d."""
    N, Do, QH, QW = qshape
    _, D, H, W, _ = src.shape
    ys = torch.arange(QH) * lat + ph[1] + org[1] + d[1]
    xs = torch.arange(QW) * lat + ph[0] + org[0] + d[0]
    zs = torch.arange(Do) * latz + phz + org[2] + d[2]
    out = torch.zeros((N, Do, QH, QW, kc), dtype=torch.float32)
    vz, vy, vx = (zs >= 0) & (zs < D), (ys >= 0) & (ys < H), (xs >= 0) & (xs < W)
    if vz.any() and vy.any() and vx.any():
        sub = src[:, zs[vz]][:, :, ys[vy]][:, :, :, xs[vx]][..., c0:c0 + kc].float()
        iz, iy, ix = vz.nonzero().flatten(), vy.nonzero().flatten(), vx.nonzero().flatten()
        out[:, iz[:, None, None], iy[None, :, None], ix[None, None, :]] = sub
    return out


def _split16(v):
    """(hi, lo) fp16 pair of an fp32 tensor: hi = fp16(v), lo = fp16(v - hi)"""
    hi = v.half()
    return hi, (v - hi.float()).half()


def _store(out, sl, v, plan, out_coff):
    """epilogue store: fp16, or the (hi, lo) halves in strict mode"""
    if plan.split_out:
        hi, lo = _split16(v)
        out[sl + (slice(0, plan.Co),)] = hi
        out[sl + (slice(plan.Co, 2 * plan.Co),)] = lo
    else:
        out[sl + (slice(out_coff, out_coff + plan.Co),)] = v.half()


def range_scale(x):
    """tpz_range_scale: (s, 1/s), s a power of two <= 1; 1 when max|x| <= 64"""
    amax = float(x.abs().max()) if x.numel() else 0.0
    s = 1.0
    if math.isfinite(amax) and amax > 64.0:
        m, e = math.frexp(np.float32(amax))
        s = math.ldexp(1.0, max(-100, 3 - e))
    return torch.tensor([s, 1.0 / s], dtype=torch.float32)


def _rs(rng):
    return (float(rng[0]), float(rng[1])) if rng is not None else (1.0, 1.0)


def tc_conv(plan, srcs, out_shape, out=None, res=None, res_org=(0, 0, 0), dot_out=None, out_coff=0, tag=None, dot_affine=None,
            rng=None):
    N, Do, Ho, Wo = out_shape
    s_, inv_s = _rs(rng)
    osc = plan.oscale if plan.oscale is not None else 1.0
    mixed = plan.lattice_z > 1 or plan.phase_sel != 0 or any(l not in (0, plan.lattice) for l in (plan.lats or [])) or any(not p for p in (plan.phases or []))
    if not mixed:
        return _tc_conv_dense(plan, srcs, out_shape, out, res, res_org, dot_out, out_coff, dot_affine, rng)
    L = plan.lattice
    phases = [plan.phase_sel - 1] if plan.phase_sel else list(range(L * L))
    assert res is None and dot_out is None
    for ph in phases:
        phy, phx = ph // L, ph % L
        QH, QW = (Ho - phy + L - 1) // L, (Wo - phx + L - 1) // L
        Lz, pz = plan.lattice_z, plan.phase_z
        QD = (Do - pz + Lz - 1) // Lz
        acc = torch.zeros((N, QD, QH, QW, plan.Co), dtype=torch.float32)
        for kb, (dx, dy, dz, c0, si) in enumerate(plan.kblocks):
            lat = plan.lats[si] or L
            p_on = plan.phases[si]
            latz = (plan.lat_zs[si] if plan.lat_zs else 0) or Lz
            A = _gather_lattice(srcs[si], plan.orgs[si], (dx, dy, dz), (N, QD, QH, QW), c0, plan.KC, lat,
                                (phx if p_on else 0, phy if p_on else 0), latz, pz if p_on else 0)
            acc += A @ plan.weights[kb].float().t()
        v = acc * osc + plan.bias * s_
        v = torch.where(v > 0, v, v * plan.neg_slope)
        _store(out, (slice(None), slice(pz, None, Lz), slice(phy, None, L), slice(phx, None, L)), v, plan, out_coff)


def _tc_conv_dense(plan, srcs, out_shape, out, res, res_org, dot_out, out_coff, dot_affine, rng=None):
    N, Do, Ho, Wo = out_shape
    s_, inv_s = _rs(rng)
    osc = plan.oscale if plan.oscale is not None else 1.0
    acc = torch.zeros((N, Do, Ho, Wo, plan.Co), dtype=torch.float32)
    for kb, (dx, dy, dz, c0, si) in enumerate(plan.kblocks):
        A = _gather(srcs[si], plan.orgs[si], (dx, dy, dz), out_shape, c0, plan.KC)
        acc += A @ plan.weights[kb].float().t()
    v = acc * osc + plan.bias * s_
    if res is not None:
        r = res[:, res_org[2]:res_org[2] + Do, res_org[1]:res_org[1] + Ho, res_org[0]:res_org[0] + Wo, :plan.Co].float()
        v = v + (r * plan.res_scale if plan.res_scale is not None else r)
    v = torch.where(v > 0, v, v * plan.neg_slope)
    if dot_out is not None:
        dv = (v * plan.dot_w).sum(-1) * inv_s + plan.dot_b
        if dot_affine is not None:
            dv = dv * dot_affine[1] + dot_affine[0]
        dot_out.copy_(dv)
    if out is not None:
        _store(out, (Ellipsis,), v, plan, out_coff)


def conv_first(x, w, bias, dil, pad, neg_slope, out_ld, rng=None, split=False):
    N, D, H, W = x.shape
    Co, kd, kh, kw = w.shape
    s_, _ = _rs(rng)
    x = x * s_
    bias = bias * s_ if bias is not None else None
    if kd > 1:
        y = F.conv3d(x[:, None], w[:, None], bias, dilation=dil, padding=pad)
    else:
        y = F.conv2d(x.reshape(N * D, 1, H, W), w, bias, dilation=dil, padding=pad)
        y = y.reshape(N, D, Co, y.shape[-2], y.shape[-1]).permute(0, 2, 1, 3, 4)
    y = torch.where(y > 0, y, y * neg_slope).permute(0, 2, 3, 4, 1)
    out = torch.zeros(tuple(y.shape[:4]) + (2 * out_ld if split else out_ld,), dtype=torch.float16)
    hi, lo = _split16(y)
    out[..., :Co] = hi
    if split:
        out[..., out_ld:out_ld + Co] = lo
    return out


def conv_first_tc(x, w_packed, bias, k, pad, neg_slope, pool=False, rng=None):
    """tpz_conv_first_tc: fp16 taps x fp16 weights, fp32 accumulate, bias + activation, fp16 NHWC out."""
    kb, cp, _ = w_packed.shape
    w = w_packed.float().permute(1, 0, 2).reshape(cp, kb * 64)[:, :k * k].reshape(cp, 1, k, k)
    s_, _ = _rs(rng)
    y = F.conv2d((x * s_).half().float()[:, None], w, bias.float() * s_, padding=pad)
    y = torch.where(y > 0, y, y * neg_slope).half().float()
    if pool:
        y = F.max_pool2d(y, 2)
    return y.permute(0, 2, 3, 1)[:, None].contiguous().half()


def im2col3d_first(x, k, ld, rng=None, split=False):
    N, D, H, W = x.shape
    p = k // 2
    xp = F.pad(x * _rs(rng)[0], (p, p, p, p, p, p))
    out = torch.zeros((N, D, H, W, 2 * ld if split else ld), dtype=torch.float16)
    t = 0
    for dz in range(k):
        for dy in range(k):
            for dx in range(k):
                hi, lo = _split16(xp[:, dz:dz + D, dy:dy + H, dx:dx + W])
                out[..., t] = hi
                if split:
                    out[..., ld + t] = lo
                t += 1
    return out


def im2col_first(x, k, pad, ld, rng=None, split=False):
    N, H, W = x.shape
    xp = F.pad(x * _rs(rng)[0], (pad, pad, pad, pad))
    Ho, Wo = H + 2 * pad - (k - 1), W + 2 * pad - (k - 1)
    out = torch.zeros((N, 1, Ho, Wo, 2 * ld if split else ld), dtype=torch.float16)
    for r in range(k):
        for s in range(k):
            hi, lo = _split16(xp[:, r:r + Ho, s:s + Wo])
            out[:, 0, :, :, r * k + s] = hi
            if split:
                out[:, 0, :, :, ld + r * k + s] = lo
    return out


def conv_last(x, c_real, w, bias, kdhw, dil, pad, stats=None, out_scale=1.0, out_shift=0.0, rng=None):
    N, D, H, W, ld = x.shape
    kd, kh, kw = kdhw
    C = w.shape[1]
    wt = w.t().reshape(1, C, kd, kh, kw)
    xi = x[..., :C].float().permute(0, 4, 1, 2, 3)          # split inputs: C = 2 * channels, weights repeated
    y = F.conv3d(xi, wt, None, dilation=dil, padding=(pad if kd > 1 else 0, pad, pad))[:, 0]
    y = (y * _rs(rng)[1] + bias) * out_scale + out_shift
    if stats is not None:
        y = y * stats[1] + stats[0]
    return y


def maxpool2(x, dims, split=False):
    if split:
        C = x.shape[-1] // 2
        v = x[..., :C].float() + x[..., C:].float()
        y = F.max_pool3d(v.permute(0, 4, 1, 2, 3), (2 if dims == 3 else 1, 2, 2)).permute(0, 2, 3, 4, 1)
        return torch.cat(_split16(y), dim=-1).contiguous()
    xi = x.float().permute(0, 4, 1, 2, 3)
    y = F.max_pool3d(xi, (2 if dims == 3 else 1, 2, 2))
    return y.permute(0, 2, 3, 4, 1).half().contiguous()


def upsample_nearest(x, size):
    N, D, H, W, C = x.shape
    Do, Ho, Wo = size
    def idx(n_in, n_out):
        scale = np.float32(n_in) / np.float32(n_out)
        return torch.tensor([min(int(math.floor(np.float32(i) * scale)), n_in - 1) for i in range(n_out)])
    return x[:, idx(D, Do)][:, :, idx(H, Ho)][:, :, :, idx(W, Wo)].contiguous()


def filter_f32(x, f, bias=0.0):
    kd, kh, kw = f.shape
    y = F.conv3d(x[:, None], f[None, None], None, padding=(kd // 2, kh // 2, kw // 2))[:, 0]
    return y + bias


def meanstd(x, unbiased):
    xd = x.double()
    return torch.stack([xd.mean(), xd.std(unbiased=unbiased)]).float()


def affine(x, stats, inverse=False, out=None):
    y = x * stats[1] + stats[0] if inverse else (x - stats[0]) / stats[1]
    if out is not None:
        out.copy_(y); return out
    return y


def gemm_f32(A, B, C):
    C.copy_((A.double() @ B.double()).float())
    return C


def gmm_sums(x, shift, sets8, work=None):
    """tpz_gmm_sums in float64 numpy."""
    import numpy as np
    xf = x.numpy().ravel()
    xc = xf.astype(np.float64) - float(shift)
    out = []
    for mode, split, mu0, mu1, var0, var1, lp0, lp1 in np.asarray(sets8, dtype=np.float64).reshape(-1, 8):
        if mode == 0:
            p0 = (xf <= np.float32(split)).astype(np.float64); p1 = 1.0 - p0; Z = np.zeros_like(xc)
        else:
            l0 = -(xc - mu0) ** 2 / 2 / var0 - 0.5 * np.log(2 * np.pi * var0) + lp0
            l1 = -(xc - mu1) ** 2 / 2 / var1 - 0.5 * np.log(2 * np.pi * var1) + lp1
            ma = np.maximum(l0, l1)
            Z = ma + np.log(np.exp(l0 - ma) + np.exp(l1 - ma))
            p0, p1 = np.exp(l0 - Z), np.exp(l1 - Z)
        out.append([Z.sum(), p0.sum(), p1.sum(), (p0 * xc).sum(), (p1 * xc).sum(), (p0 * xc * xc).sum(), (p1 * xc * xc).sum()])
    return np.array(out)


def select_hist(x, level, prefixes=()):
    import numpy as np
    b = x.numpy().ravel().view(np.uint32).astype(np.uint64)
    key = np.where(b & 0x80000000, (~b) & 0xFFFFFFFF, b | 0x80000000).astype(np.uint64)
    if level == 0:
        return np.bincount((key >> 20).astype(np.int64), minlength=4096).astype(np.int64)
    rows = []
    for p in prefixes:
        if level == 1:
            sel = key[(key >> 20) == p]; rows.append(np.bincount(((sel >> 8) & 0xFFF).astype(np.int64), minlength=4096))
        else:
            sel = key[(key >> 8) == p]; rows.append(np.bincount((sel & 0xFF).astype(np.int64), minlength=256))
    return np.stack(rows).astype(np.int64)


def to_device(t):
    return t


@contextlib.contextmanager
def patched():
    names = ['tc_conv', 'range_scale', 'conv_first', 'conv_first_tc', 'im2col_first', 'im2col3d_first', 'filter_f32', 'conv_last', 'maxpool2', 'upsample_nearest', 'meanstd', 'affine',
             'gemm_f32', 'gmm_sums', 'select_hist', 'to_device']
    from topaz_b200 import engine
    saved = {n: getattr(ops, n) for n in names}
    saved['require_cuda'] = ops.require_cuda
    dense_engine = engine.DENSE_ENGINE
    try:
        for n in names:
            setattr(ops, n, globals()[n])
        ops.require_cuda = lambda t, what: None
        engine.DENSE_ENGINE = 'py'      # the simulation replaces the layer-level wrappers; the model-level C ABI needs the GPU
        yield
    finally:
        engine.DENSE_ENGINE = dense_engine
        for n, f in saved.items():
            setattr(ops, n, f)


# ------------------------------------------------------------------------------------------------
# training kernels (topaz_b200.train_engine wrappers) simulated with torch CPU ops
# ------------------------------------------------------------------------------------------------
def _nchw(t):
    return t.permute(0, 3, 1, 2)


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _shift_in(x, org, Ho, Wo, k, dil, stride):
    """input window so that tap 0 of output pixel 0 sits at `org` (org >= 0 in the training nets)."""
    need_h = (Ho - 1) * stride + (k - 1) * dil + 1
    need_w = (Wo - 1) * stride + (k - 1) * dil + 1
    return x[:, org:org + need_h, org:org + need_w]


def t_conv_fwd(x, w, b, stride, dil, org, Ho, Wo, relu, res=None, res_org=0, res_stride=1):
    k = w.shape[-1]
    xin = _shift_in(x, org, Ho, Wo, k, dil, stride)
    y = F.conv2d(_nchw(xin), w.detach(), b.detach() if b is not None else None, stride=stride, dilation=dil)
    y = _nhwc(y)
    if res is not None:
        y = y + res[:, res_org:res_org + (Ho - 1) * res_stride + 1:res_stride, res_org:res_org + (Wo - 1) * res_stride + 1:res_stride]
    return torch.relu(y) if relu else y


def t_conv_dgrad(dy, w, stride, dil, org, H, W, mask=None, accumulate=False, out=None, res=None, res_org=0):
    N, Ho, Wo, Co = dy.shape
    k = w.shape[-1]
    need_h = (Ho - 1) * stride + (k - 1) * dil + 1
    need_w = (Wo - 1) * stride + (k - 1) * dil + 1
    gi = torch.nn.grad.conv2d_input((N, w.shape[1], need_h, need_w), w.detach(), _nchw(dy).contiguous(), stride=stride, dilation=dil)
    dx = torch.zeros((N, H, W, w.shape[1]))
    dx[:, org:org + need_h, org:org + need_w] = _nhwc(gi)
    if accumulate:
        dx = dx + out
    if res is not None:
        dx[:, res_org:res_org + res.shape[1], res_org:res_org + res.shape[2]] += res
    if mask is not None:
        dx = torch.where(mask > 0, dx, torch.zeros_like(dx))
    if out is not None:
        out.copy_(dx); return out
    return dx


def t_conv_wgrad(x, dy, w_grad, b_grad, stride, dil, org):
    N, Ho, Wo, Co = dy.shape
    k = w_grad.shape[-1]
    xin = _shift_in(x, org, Ho, Wo, k, dil, stride)
    gw = torch.nn.grad.conv2d_weight(_nchw(xin).contiguous(), tuple(w_grad.shape), _nchw(dy).contiguous(), stride=stride, dilation=dil)
    w_grad.add_(gw)
    if b_grad is not None:
        b_grad.add_(dy.sum((0, 1, 2)))


def t_cls_bwd(x, g, w, w_grad, b_grad, masked):
    t_conv_wgrad(x, g, w_grad, b_grad, 1, 1, 0)
    return t_conv_dgrad(g, w, 1, 1, 0, x.shape[1], x.shape[2], mask=x if masked else None)


def t_relu_bwd(dy, y):
    dy.mul_((y > 0).float())


def t_crop_add(dx, g, org, stride):
    Ho, Wo = g.shape[1], g.shape[2]
    dx[:, org:org + (Ho - 1) * stride + 1:stride, org:org + (Wo - 1) * stride + 1:stride] += g


def t_bn_stats(x, sums):
    C = x.shape[-1]
    xd = x.double().reshape(-1, C)
    sums[:C] += xd.sum(0)
    sums[C:] += (xd * xd).sum(0)


def t_bn_fwd(x, sums, count, gamma, beta, eps, momentum, running_mean, running_var, relu, save):
    C = x.shape[-1]
    if sums is not None:
        mean = sums[:C] / count
        var = (sums[C:] / count - mean * mean).clamp_min(0)
        invstd = 1.0 / torch.sqrt(var + eps)
        save[:C] = mean.float()
        save[C:] = invstd.float()
        if running_mean is not None:
            running_mean.mul_(1 - momentum).add_(momentum * mean.float())
            running_var.mul_(1 - momentum).add_(momentum * (var * (count / (count - 1.0))).float())
    y = (x - save[:C]) * (save[C:] * gamma.detach()) + beta.detach()
    return torch.relu(y) if relu else y


def t_bn_bwd_reduce(g, x, save, sums):
    C = x.shape[-1]
    xhat = ((x - save[:C]) * save[C:]).double().reshape(-1, C)
    gd = g.double().reshape(-1, C)
    sums[:C] += gd.sum(0)
    sums[C:] += (gd * xhat).sum(0)


def t_bn_bwd(g, x, save, sums, count, gamma, local_sums, dgamma, dbeta):
    C = x.shape[-1]
    xhat = (x - save[:C]) * save[C:]
    a, b = (sums[:C] / count).float(), (sums[C:] / count).float()
    dx = (save[C:] * gamma.detach()) * (g - a - xhat * b)
    dbeta.add_(local_sums[:C].float())
    dgamma.add_(local_sums[C:].float())
    g.copy_(dx)


def _slope(act):
    return act.weight.detach().reshape(()) if isinstance(act, torch.nn.PReLU) else torch.tensor(float(act.negative_slope))


def t_act_fwd(v, act):
    return torch.where(v > 0, v, _slope(act) * v)


def t_act_bwd(g, v, act):
    neg = ~(v > 0)
    if isinstance(act, torch.nn.PReLU):
        act.weight.grad.add_((g[neg].double() * v[neg].double()).sum().float())
    g.copy_(torch.where(neg, _slope(act) * g, g))
    return g


DROPOUT_REPLAY = []          # tests may queue keep-masks (NHWC, bool) recorded from the reference's nn.Dropout layers


def t_dropout_fwd(x, p):
    keep = DROPOUT_REPLAY.pop(0) if DROPOUT_REPLAY else (torch.rand_like(x) > p)
    return x * (keep.to(x.dtype) / (1.0 - p)), keep.to(torch.uint8)


def t_dropout_bwd(g, mask, p):
    g.copy_(g * (mask.to(g.dtype) / (1.0 - p)))
    return g


def t_ge_loss_grad(scores, labels, pi, slack, lo, hi, dscore, out5):
    from oracle import topaz_oracle as O
    s = scores.detach().clone().requires_grad_(True)
    cls, ge, loss = O.ge_binomial_loss(s, labels, pi, slack)
    loss.backward()
    dscore.copy_(s.grad[lo:hi].float())
    prec, tpr, fpr = O.ge_binomial_metrics(scores, labels)
    out5.copy_(torch.tensor([cls.item(), ge.item(), prec, tpr, fpr]))


def t_pu_objective(scores, labels, mode, pi, slack, momentum, aux_in, lo, hi, dscore, out6):
    from oracle import topaz_oracle as O
    s = scores.detach().clone().requires_grad_(True)
    name = ['PN', 'GE_KL', 'PU'][mode]
    rep, ge, back, new_run = O.pu_objective_loss(name, s, labels, (pi if pi > 0 else None) if mode == 0 else pi, slack, momentum,
                                                 aux_in, aux_in)
    back.backward()
    dscore.copy_(s.grad[lo:hi].float())
    prec, tpr, fpr = O.ge_binomial_metrics(scores, labels)
    out6.copy_(torch.tensor([rep.item(), ge.item() if ge is not None else 0.0, prec, tpr, fpr, new_run if new_run is not None else 0.0]))


def t_adam_step(fp, lr, b1, b2, eps, l2):
    fp.step += 1
    g = fp.flat_g + l2 * fp.flat_p
    fp.flat_m.mul_(b1).add_(g, alpha=1 - b1)
    fp.flat_v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** fp.step, 1 - b2 ** fp.step
    fp.flat_p.sub_((lr / bc1) * fp.flat_m / (fp.flat_v.sqrt() / math.sqrt(bc2) + eps))
    fp.flat_g.zero_()


@contextlib.contextmanager
def patched_training():
    from topaz_b200 import train_engine as T
    names = {'_conv_fwd': t_conv_fwd, '_conv_dgrad': t_conv_dgrad, '_conv_wgrad': t_conv_wgrad, '_cls_bwd': t_cls_bwd, '_relu_bwd': t_relu_bwd,
             '_crop_add': t_crop_add, '_bn_stats': t_bn_stats, '_bn_fwd': t_bn_fwd, '_bn_bwd_reduce': t_bn_bwd_reduce,
             '_bn_bwd': t_bn_bwd, '_act_fwd': t_act_fwd, '_act_bwd': t_act_bwd, '_dropout_fwd': t_dropout_fwd, '_dropout_bwd': t_dropout_bwd,
             'ge_loss_grad': t_ge_loss_grad, 'pu_objective_loss_grad': t_pu_objective, 'adam_step': t_adam_step,
             'read_back': lambda d, h: d.tolist(), '_repack': lambda fp, force=False: None}
    saved = {n: getattr(T, n) for n in names}
    try:
        for n, f in names.items():
            setattr(T, n, f)
        with patched():
            yield
    finally:
        for n, f in saved.items():
            setattr(T, n, f)
