"""CPU oracle for the Topaz dense-CNN hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional fp32 CPU restatement (torch CPU ops + numpy) of the
reference algorithms on the hot path.  Nothing in ``topaz_b200/`` may import it;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs do, and only as the checker / the timed CPU baseline.

Parity pin: the oracle is checked against golden vectors produced by the REAL
reference (imported from /root/reference by ``tools/make_goldens.py``) in
``tests/test_oracle_golden.py``.  The reference's own test-suite pins no numeric
result on this path, so those goldens + the packaged
pretrained weights are the pin.  Training-mode BatchNorm, the PReLU extractors and dropout are pinned the same way
(``tools/make_goldens_bn.py``: 2-3 GE_binomial steps of the real reference; for dropout, with the keep-masks its
nn.Dropout layers drew recorded by hooks, since torch's mask stream cannot be reproduced outside torch).

Every function cites the reference file:line it restates.  Weights are passed
as a plain ``dict[str, np.ndarray | torch.Tensor]`` keyed by the reference's
``state_dict`` names.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def _t(a) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        if a.requires_grad:          # training oracle: keep the autograd graph
            return a
        return a.detach().to(torch.float32).cpu()
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.float32)


def _conv(x, w, b=None, stride=1, dilation=1, padding=0):
    if w.dim() == 5:
        return F.conv3d(x, w, b, stride=stride, dilation=dilation, padding=padding)
    return F.conv2d(x, w, b, stride=stride, dilation=dilation, padding=padding)


# --------------------------------------------------------------------------------------
# Architecture descriptions (shared vocabulary with the product, restated independently)
# --------------------------------------------------------------------------------------

def resnet_spec(kind: str, units: int) -> List[dict]:
    """Block list of ResNet8 / ResNet16 (reference: topaz/model/features/resnet.py:280-339).

    Each entry: {'type': 'conv'|'resid', 'k', 'dil', 'stride', 'nin', 'nout', ...} in the
    *unfilled* (training) geometry; pooling=None so stride=2 on the strided blocks.
    """
    u0, u1, u2 = units, 2 * units, 4 * units
    if kind == 'resnet8':
        return [
            dict(type='conv', nin=1, nout=u0, k=7, dil=1, stride=2),
            dict(type='resid', nin=u0, nhid=u0, nout=u0, dil=2, stride=1),
            dict(type='resid', nin=u0, nhid=u0, nout=u1, dil=2, stride=2),
            dict(type='resid', nin=u1, nhid=u1, nout=u1, dil=2, stride=1),
            dict(type='conv', nin=u1, nout=u2, k=5, dil=1, stride=1),
        ]
    if kind == 'resnet16':
        return [
            dict(type='conv', nin=1, nout=u0, k=7, dil=1, stride=1),
            dict(type='resid', nin=u0, nhid=u0, nout=u0, dil=1, stride=2),
            dict(type='resid', nin=u0, nhid=u0, nout=u0, dil=1, stride=1),
            dict(type='resid', nin=u0, nhid=u0, nout=u0, dil=1, stride=1),
            dict(type='resid', nin=u0, nhid=u0, nout=u0, dil=1, stride=1),
            dict(type='resid', nin=u0, nhid=u0, nout=u1, dil=1, stride=2),
            dict(type='resid', nin=u1, nhid=u1, nout=u1, dil=1, stride=1),
            dict(type='resid', nin=u1, nhid=u1, nout=u1, dil=1, stride=1),
            dict(type='conv', nin=u1, nout=u2, k=5, dil=1, stride=1),
        ]
    raise ValueError(kind)


def resnet_width(spec: Sequence[dict]) -> int:
    """Receptive field (reference: topaz/model/utils.py:39-68 applied to the module list)."""
    out = 1
    for blk in reversed(spec):
        if blk['type'] == 'conv':
            out = (out - 1) * blk['stride'] + 1 + (blk['k'] - 1) * blk['dil']
        else:  # ResidA exposes kernel_size = 2*dilation+3, dilation attr = 1 (resnet.py:135-137)
            out = (out - 1) * blk['stride'] + 1 + (2 * blk['dil'] + 3 - 1)
    return out


def _bn_eval(x, sd, prefix, eps=1e-5):
    g, b = _t(sd[prefix + '.weight']), _t(sd[prefix + '.bias'])
    m, v = _t(sd[prefix + '.running_mean']), _t(sd[prefix + '.running_var'])
    shape = (1, -1) + (1,) * (x.dim() - 2)
    return (x - m.view(shape)) / torch.sqrt(v.view(shape) + eps) * g.view(shape) + b.view(shape)


def _bn_train(x, sd, prefix, eps=1e-5, momentum=0.1, running=None):
    """nn.BatchNorm2d in training mode (resnet.py:68-70,134-141): biased batch variance for the normalisation,
    running_mean/var <- (1-momentum)*running + momentum*(batch mean, UNBIASED batch variance).  `running` (a dict)
    receives the updated buffers."""
    g, b = _t(sd[prefix + '.weight']), _t(sd[prefix + '.bias'])
    dims = [0] + list(range(2, x.dim()))
    shape = (1, -1) + (1,) * (x.dim() - 2)
    mean = x.mean(dims)
    var = x.var(dims, unbiased=False)
    if running is not None:
        n = x.numel() // x.shape[1]
        with torch.no_grad():
            running[prefix + '.running_mean'] = (1 - momentum) * _t(sd[prefix + '.running_mean']).detach() + momentum * mean.detach()
            running[prefix + '.running_var'] = (1 - momentum) * _t(sd[prefix + '.running_var']).detach() + \
                momentum * var.detach() * (n / (n - 1.0))
    return (x - mean.view(shape)) / torch.sqrt(var.view(shape) + eps) * g.view(shape) + b.view(shape)


PREACT_LOG = None      # tests set this to a list to record every pre-activation (margin-from-zero checks of gradient tests)


def _act(y, relu_masks):
    """ReLU; with `relu_masks` (an iterator of 0/1 tensors in layer order) the mask is imposed instead of derived from
    y, so that a gradient check is not at the mercy of activations within rounding noise of zero."""
    if PREACT_LOG is not None:
        PREACT_LOG.append(y.detach())
    if relu_masks is None:
        return F.relu(y)
    return y * next(relu_masks).to(y.dtype)


def _dropout(x, p, dropout_masks):
    """nn.Dropout(p) in training with the keep-mask imposed (an iterator of 0/1 tensors shaped like x): x * mask / (1-p).
    torch's own mask stream cannot be reproduced outside torch, so the masks always come from the implementation under test;
    the arithmetic is pinned against F.dropout in tests/test_oracle_golden.py."""
    return x * (next(dropout_masks).to(x.dtype) / (1.0 - p))        # ATen: input * (bernoulli(1-p) / (1-p))


RESNET_DROPOUT_AFTER = {'resnet8': (0, 2, 4), 'resnet16': (1, 5, 8)}     # spec indices followed by nn.Dropout (resnet.py:294-336)


def resnet_features(sd: Dict, x: torch.Tensor, kind: str, units: int, filled: bool,
                    bn: bool = False, prefix: str = 'features.features.', bn_train: bool = False,
                    running: Optional[Dict] = None, relu_masks=None, dropout: float = 0.0,
                    dropout_masks=None) -> torch.Tensor:
    """ResNet.forward (resnet.py:243-251) with fill() semantics (resnet.py:87-92,153-164,227-232).

    filled=True: input padded by width//2 once, every stride -> 1, dilations multiplied by the
    cumulative stride.  filled=False: the strided training geometry.  BN (eval mode only here)
    sits after conv for BasicConv (resnet.py:101-105) and after the residual add for ResidA
    (resnet.py:199-202).
    """
    spec = resnet_spec(kind, units)
    if x.dim() < 4:
        x = x.unsqueeze(1)
    if filled:
        p = resnet_width(spec) // 2
        x = F.pad(x, (p, p, p, p))
    cum = 1
    drop_after = RESNET_DROPOUT_AFTER[kind] if dropout > 0 else ()
    mod_idx = 0                      # index in the nn.Sequential: Dropout modules occupy slots too (state_dict keys shift)
    for i, blk in enumerate(spec):
        pre = f'{prefix}{mod_idx}.'
        mod_idx += 2 if i in drop_after else 1
        if blk['type'] == 'conv':
            w = _t(sd[pre + 'conv.weight'])
            b = _t(sd[pre + 'conv.bias']) if (pre + 'conv.bias') in sd else None
            if filled:
                y = _conv(x, w, b, stride=1, dilation=blk['dil'] * cum)
                cum *= blk['stride']
            else:
                y = _conv(x, w, b, stride=blk['stride'], dilation=blk['dil'])
            if bn:
                y = _bn_train(y, sd, pre + 'bn', running=running) if bn_train else _bn_eval(y, sd, pre + 'bn')
            x = _act(y, relu_masks)
            if i in drop_after and dropout_masks is not None:
                x = _dropout(x, dropout, dropout_masks)
        else:
            w0 = _t(sd[pre + 'conv0.weight'])
            b0 = _t(sd[pre + 'conv0.bias']) if (pre + 'conv0.bias') in sd else None
            w1 = _t(sd[pre + 'conv1.weight'])
            b1 = _t(sd[pre + 'conv1.bias']) if (pre + 'conv1.bias') in sd else None
            if filled:
                d0, d1, s = cum, blk['dil'] * cum, 1
            else:
                d0, d1, s = 1, blk['dil'], blk['stride']
            h = _conv(x, w0, b0, dilation=d0)
            if bn:
                h = _bn_train(h, sd, pre + 'bn0', running=running) if bn_train else _bn_eval(h, sd, pre + 'bn0')
            h = _act(h, relu_masks)
            y = _conv(h, w1, b1, stride=s, dilation=d1)
            edge = d0 + d1
            xs = x[:, :, edge:-edge, edge:-edge]
            if (pre + 'proj.weight') in sd:
                xs = _conv(xs, _t(sd[pre + 'proj.weight']), None, stride=s)
            elif s > 1:
                xs = xs[..., ::s, ::s]
            y = y + xs
            if bn:
                y = _bn_train(y, sd, pre + 'bn1', running=running) if bn_train else _bn_eval(y, sd, pre + 'bn1')
            x = _act(y, relu_masks)
            if i in drop_after and dropout_masks is not None:
                x = _dropout(x, dropout, dropout_masks)
            if filled:
                cum *= blk['stride']
    return x


def basicconv_layers(sizes: Sequence[int], units: int, unit_scaling: int) -> List[dict]:
    """conv31/63/127 stacks (reference: topaz/model/features/basic.py:12-78, factory.py:15-25)."""
    out, nin = [], 1
    for k in sizes[:-1]:
        out.append(dict(nin=nin, nout=units, k=k, stride=2))
        nin, units = units, units * unit_scaling
    out.append(dict(nin=nin, nout=units, k=sizes[-1], stride=1))
    return out


def basicconv_width(layers: Sequence[dict]) -> int:
    out = 1
    for l in reversed(layers):
        out = (out - 1) * l['stride'] + 1 + (l['k'] - 1)
    return out


def _prelu(y, a, act_masks):
    """PReLU with one slope; with `act_masks` (iterator of boolean "v > 0" tensors) the branch is imposed, see _act."""
    if act_masks is None:
        return F.prelu(y, a)
    return torch.where(next(act_masks), y, a.reshape(()) * y)


def basicconv_features(sd: Dict, x: torch.Tensor, sizes: Sequence[int], units: int, unit_scaling: int,
                       filled: bool, bn: bool = True, prefix: str = 'features.features.', bn_train: bool = False,
                       running: Optional[Dict] = None, act_masks=None, dropout: float = 0.0,
                       dropout_masks=None) -> torch.Tensor:
    """BasicConv.forward + fill (basic.py:81-111): conv -> (BN) -> PReLU(1 slope) per layer;
    filled => stride 1, dilation = cumulative stride, one pad of width//2 at the input.
    bn_train: BatchNorm in training mode (minibatch statistics; `running` receives the updated buffers)."""
    layers = basicconv_layers(sizes, units, unit_scaling)
    if x.dim() < 4:
        x = x.unsqueeze(1)
    if filled:
        p = basicconv_width(layers) // 2
        x = F.pad(x, (p, p, p, p))
    idx, cum = 0, 1
    for l in layers:
        w = _t(sd[f'{prefix}{idx}.weight'])
        b = _t(sd[f'{prefix}{idx}.bias']) if f'{prefix}{idx}.bias' in sd else None
        if filled:
            y = _conv(x, w, b, stride=1, dilation=cum)
            cum *= l['stride']
        else:
            y = _conv(x, w, b, stride=l['stride'])
        idx += 1
        if bn:
            y = _bn_train(y, sd, f'{prefix}{idx}', running=running) if bn_train else _bn_eval(y, sd, f'{prefix}{idx}')
            idx += 1
        a = _t(sd[f'{prefix}{idx}.weight'])
        x = _prelu(y, a, act_masks)
        idx += 1
        if dropout > 0:              # nn.Dropout follows every activation (basic.py:58-59,71-72) and occupies a module slot
            idx += 1
            if dropout_masks is not None:
                x = _dropout(x, dropout, dropout_masks)
    return x


def classifier_forward(sd: Dict, x, arch: str, units: int, filled: bool, bn: bool = False,
                       unit_scaling: int = 1) -> torch.Tensor:
    """LinearClassifier.forward (classifier.py:48-66): features then 1x1 conv to one logit."""
    x = _t(x)
    if arch in ('resnet8', 'resnet16'):
        z = resnet_features(sd, x, arch, units, filled, bn)
    else:
        sizes = {'conv31': [7, 5, 5], 'conv63': [7, 5, 5, 5], 'conv127': [7, 5, 5, 5, 5]}[arch]
        z = basicconv_features(sd, x, sizes, units, unit_scaling, filled, bn, prefix='features.features.')
    return _conv(z, _t(sd['classifier.weight']), _t(sd['classifier.bias']))


# --------------------------------------------------------------------------------------
# U-Net denoisers
# --------------------------------------------------------------------------------------

def unet_forward(sd: Dict, x) -> torch.Tensor:
    """UDenoiseNet.forward / UDenoiseNetSmall.forward / UDenoiseNet3D.forward (denoising/models.py:130-175,
    221-244, 508-564).

    Dimensionality, depth (6 encoder stages; 4 for UDenoiseNetSmall), base/top widths are inferred from the
    state dict.  Encoder: conv(same pad) + LeakyReLU(0.1) + MaxPool(2) (last stage without pool).  Decoder:
    nearest-upsample to the skip's size, concat [upsampled, skip], two conv+LeakyReLU; dec1: three convs, last
    one linear.
    """
    x = _t(x)
    nd = _t(sd['enc1.0.weight']).dim() - 2
    pool = F.max_pool3d if nd == 3 else F.max_pool2d

    def cv(h, name):
        w = _t(sd[name + '.weight'])
        return _conv(h, w, _t(sd[name + '.bias']), padding=w.shape[-1] // 2)

    depth = max(i for i in range(1, 10) if f'enc{i}.0.weight' in sd)
    skips = [x]
    h = x
    for i in range(1, depth):
        h = pool(F.leaky_relu(cv(h, f'enc{i}.0'), 0.1), 2)
        skips.append(h)
    h = F.leaky_relu(cv(h, f'enc{depth}.0'), 0.1)
    # skips = [x, p1, ..., p_{depth-1}]; dec_{depth-1} joins p_{depth-2}, ..., dec1 joins x
    for lvl in range(depth - 1, 0, -1):
        skip = skips[lvl - 1]
        h = F.interpolate(h, size=tuple(skip.shape[2:]), mode='nearest')
        h = torch.cat([h, skip], 1)
        h = F.leaky_relu(cv(h, f'dec{lvl}.0'), 0.1)
        h = cv(h, f'dec{lvl}.2')
        if lvl > 1:
            h = F.leaky_relu(h, 0.1)
        else:
            h = cv(F.leaky_relu(h, 0.1), 'dec1.4')
    return h


def denoise_call(sd: Dict, x, dims: int = 2) -> np.ndarray:
    """Denoise._denoise (denoise.py:274-296): mean / UNBIASED std normalise over the whole call
    input, add batch/channel dims, forward, squeeze, de-normalise."""
    x = _t(x)
    mu, std = x.mean(), x.std()
    inp = (x - mu) / std
    if inp.dim() == dims:
        inp = inp[None, None]
    elif inp.dim() == dims + 1:
        inp = inp.unsqueeze(1)
    pred = unet_forward(sd, inp).squeeze()
    return (pred * std + mu).numpy()


def denoise_patches(sd: Dict, x: np.ndarray, patch_size: int, padding: int) -> np.ndarray:
    """Denoise.denoise_patches (denoise.py:299-324)."""
    y = np.zeros_like(x)
    H, W = x.shape
    for i in range(0, H, patch_size):
        for j in range(0, W, patch_size):
            si, ei = max(0, i - padding), min(H, i + patch_size + padding)
            sj, ej = max(0, j - padding), min(W, j + patch_size + padding)
            yij = denoise_call(sd, x[si:ei, sj:ej])
            oi, oj = i - si, j - sj
            y[i:i + patch_size, j:j + patch_size] = yij[oi:oi + patch_size, oj:oj + patch_size]
    return y


def denoise(sd: Dict, x: np.ndarray, patch_size: int = -1, padding: int = 128) -> np.ndarray:
    """Denoise.denoise (denoise.py:327-332)."""
    s = patch_size + padding
    use_patch = (patch_size > 0) and (s < x.shape[0] or s < x.shape[1])
    return denoise_patches(sd, x, patch_size, padding) if use_patch else denoise_call(sd, x)


def patch3d(tomo: np.ndarray, idx: int, patch_size: int, padding: int):
    """PatchDataset.__getitem__ (denoising/datasets.py:426-468): zero-padded (p+2*pad)^3 crop."""
    shape = tuple(int(math.ceil(n / patch_size)) for n in tomo.shape)
    i, j, k = np.unravel_index(idx, shape)
    i, j, k = int(i) * patch_size, int(j) * patch_size, int(k) * patch_size
    d = patch_size + 2 * padding
    x = np.zeros((d, d, d), dtype=np.float32)
    si, ei = max(0, i - padding), min(tomo.shape[0], i + patch_size + padding)
    sj, ej = max(0, j - padding), min(tomo.shape[1], j + patch_size + padding)
    sk, ek = max(0, k - padding), min(tomo.shape[2], k + patch_size + padding)
    sic, sjc, skc = padding - i + si, padding - j + sj, padding - k + sk
    x[sic:sic + ei - si, sjc:sjc + ej - sj, skc:skc + ek - sk] = tomo[si:ei, sj:ej, sk:ek]
    return (i, j, k), x, int(np.prod(shape))


def denoise3d(sd: Dict, tomo: np.ndarray, patch_size: int = 96, padding: int = 48) -> np.ndarray:
    """Denoise3D.denoise (denoise.py:336-377): global numpy mu/std (population), patches normalised
    with them, then _denoise normalises AGAIN per batch (batch_size 1), result *std+mu, centre pasted."""
    out = np.zeros_like(tomo)
    mu, std = tomo.mean(), tomo.std()
    if patch_size < 1:
        out[:] = denoise_call(sd, tomo, dims=3)
        return out
    _, _, n = patch3d(tomo, 0, patch_size, padding)
    for p in range(n):
        (i, j, k), x, _ = patch3d(tomo, p, patch_size, padding)
        xb = torch.from_numpy(x)[None]              # DataLoader(batch_size=1) adds a batch dim
        y = denoise_call(sd, (xb - mu) / std, dims=3) * std + mu
        pz, py, px = out[i:i + patch_size, j:j + patch_size, k:k + patch_size].shape
        out[i:i + patch_size, j:j + patch_size, k:k + patch_size] = \
            y[padding:padding + pz, padding:padding + py, padding:padding + px]
    return out


def affine_normalize(x: np.ndarray):
    """stats.normalize(method='affine') (stats.py:36-46)."""
    mu, std = float(x.mean()), float(x.std())
    return ((x - mu) / std).astype(np.float32), mu, std


def gaussian_kernel(sigma: float, scale: float = 5, dims: int = 2) -> np.ndarray:
    """filters.gaussian_filter + GaussianDenoise.__init__ (filters.py:6-19, 55-59)."""
    width = 1 + 2 * int(np.ceil(sigma * scale))
    r = np.arange(-(width // 2), width // 2 + 1)
    g = np.meshgrid(*([r] * dims))
    d = sum(a ** 2 for a in g)
    f = np.exp(-0.5 * d / sigma ** 2)
    return (f / f.sum()).astype(np.float32)


def gaussian_denoise(x: np.ndarray, sigma: float, scale: float = 5) -> np.ndarray:
    """GaussianDenoise.apply (filters.py:62-79): 1->1 same-pad conv with the normalised Gaussian."""
    k = torch.from_numpy(gaussian_kernel(sigma, scale, x.ndim))[None, None]
    xt = torch.from_numpy(x)[None, None]
    return _conv(xt, k, torch.zeros(1), padding=k.shape[-1] // 2).squeeze().numpy()


# --------------------------------------------------------------------------------------
# GE-binomial training step
# --------------------------------------------------------------------------------------

def log_binom_pmf(N: int, pi: float) -> np.ndarray:
    """scipy.stats.binom.logpmf(arange(N+1), N, pi) cast to fp32 (methods.py:124-125), restated with
    lgamma so the oracle does not need scipy."""
    k = np.arange(N + 1, dtype=np.float64)
    lg = np.vectorize(math.lgamma)
    comb = lg(N + 1.0) - lg(k + 1.0) - lg(N - k + 1.0)
    return (comb + k * math.log(pi) + (N - k) * math.log1p(-pi)).astype(np.float32)


def ge_binomial_loss(score: torch.Tensor, Y: torch.Tensor, pi: float, slack: float = 1.0):
    """Loss part of GE_binomial.step (methods.py:103-136, entropy_penalty = autoencoder = 0).
    score: [B] fp32 (requires_grad ok), Y: [B] float64 labels."""
    sel1 = (Y == 1)
    cls = F.binary_cross_entropy_with_logits(score[sel1].double(), Y[sel1].double())
    sel0 = (Y == 0)
    N = int(sel0.sum().item())
    p = torch.sigmoid(score[sel0])
    q_mu = p.sum()
    q_var = torch.sum(p * (1 - p))
    k = torch.arange(0, N + 1).float()
    q = F.softmax(-0.5 * (q_mu - k) ** 2 / (q_var + 1e-10), dim=0)
    ge = -torch.sum(torch.from_numpy(log_binom_pmf(N, pi)) * q)
    return cls, ge, cls + slack * ge


def pu_objective_loss(mode: str, score: torch.Tensor, Y: torch.Tensor, pi, slack: float = 1.0, momentum: float = 1.0,
                      running: float = None, beta: float = 0.0):
    """Loss parts of PN.step (methods.py:42-53), GE_KL.step (:189-214, entropy_penalty=0) and PU.step (:282-298).
    Returns (reported_loss, ge_penalty or None, backprop_loss, new_running_expectation or None)."""
    bce = lambda s, t: F.binary_cross_entropy_with_logits(s.double(), t.double())
    pos, neg = (Y == 1), (Y == 0)
    if mode == 'PN':
        if pi is not None:
            loss = bce(score[pos], Y[pos]) * pi + bce(score[neg], Y[neg]) * (1 - pi)
        else:
            loss = bce(score, Y)
        return loss, None, loss, None
    if mode == 'GE_KL':
        cls = bce(score[pos], Y[pos])
        p_hat = torch.sigmoid(score[neg]).mean()
        new_run = None
        if momentum < 1:
            p_hat = momentum * p_hat + (1 - momentum) * running
            new_run = p_hat.item()
        entropy = pi * np.log(pi) + (1 - pi) * np.log1p(-pi)
        ge = (-torch.log(p_hat) * pi - torch.log1p(-p_hat) * (1 - pi) + entropy) * slack / momentum
        return cls, ge, cls + ge, new_run
    if mode == 'PU':
        loss_pp = bce(score[pos], Y[pos]); loss_pn = bce(score[pos], 0 * Y[pos]); loss_un = bce(score[neg], Y[neg])
        loss_u = loss_un - loss_pn * pi
        if loss_u.item() < -beta:
            return loss_pp * pi + (-beta), None, -loss_u, None
        loss = loss_pp * pi + loss_u
        return loss, None, loss, None
    raise ValueError(mode)


def ge_binomial_metrics(score: torch.Tensor, Y: torch.Tensor):
    """precision / tpr / fpr (methods.py:148-151)."""
    p = torch.sigmoid(score.detach())
    return (p[Y == 1].sum().item() / p.sum().item(), p[Y == 1].mean().item(), p[Y == 0].mean().item())


def adam_update(p, g, m, v, step, lr=2e-4, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults as used at training.py:355-356 (no weight decay / amsgrad)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * m / denom, m, v


def ge_binomial_steps(sd: Dict, Xs: Sequence[np.ndarray], Ys: Sequence[np.ndarray], arch: str, units: int,
                      pi: float, slack: float = 1.0, l2: float = 0.0, lr: float = 2e-4, bn: bool = False,
                      unit_scaling: int = 1):
    """Run len(Xs) GE_binomial.step calls (methods.py:98-165) with Adam on CPU; bn=True: a BatchNorm model in train()
    mode (minibatch statistics, running-buffer updates).
    Returns (list of 5-tuples, list of per-step grads dict, final state dict)."""
    def is_buf(k):
        return k.endswith(('running_mean', 'running_var', 'num_batches_tracked'))
    params = {k: _t(v).clone().requires_grad_(True) for k, v in sd.items() if not is_buf(k)}
    bufs = {k: torch.as_tensor(np.asarray(v)).clone() for k, v in sd.items() if is_buf(k)}
    m = {k: torch.zeros_like(v) for k, v in params.items()}
    v2 = {k: torch.zeros_like(v) for k, v in params.items()}
    outs, grads = [], []
    for t, (X, Y) in enumerate(zip(Xs, Ys), 1):
        Yt = torch.from_numpy(np.asarray(Y, dtype=np.float64))
        running = {}
        score = classifier_forward_grad({**params, **bufs}, torch.from_numpy(X), arch, units, bn=bn, running=running,
                                        unit_scaling=unit_scaling).view(-1)
        bufs.update(running)
        for k in bufs:
            if k.endswith('num_batches_tracked'):
                bufs[k] = bufs[k] + 1
        cls, ge, loss = ge_binomial_loss(score, Yt, pi, slack)
        for p_ in params.values():
            p_.grad = None
        loss.backward()
        prec, tpr, fpr = ge_binomial_metrics(score, Yt)
        if l2 > 0:
            r = 0.5 * l2 * sum(torch.sum(w ** 2) for w in params.values())
            r.backward()
        g = {k: p_.grad.detach().clone() for k, p_ in params.items()}
        grads.append({k: a.numpy() for k, a in g.items()})
        with torch.no_grad():
            for k in params:
                newp, m[k], v2[k] = adam_update(params[k].detach(), g[k], m[k], v2[k], t, lr)
                params[k].copy_(newp)
        outs.append((cls.item(), ge.item(), prec, tpr, fpr))
    final = {k: p_.detach().numpy() for k, p_ in params.items()}
    final.update({k: b.numpy() for k, b in bufs.items()})
    return outs, grads, final


def classifier_forward_grad(params: Dict, x: torch.Tensor, arch: str, units: int, bn: bool = False,
                            running: Optional[Dict] = None, relu_masks=None, unit_scaling: int = 1,
                            dropout: float = 0.0, dropout_masks=None) -> torch.Tensor:
    """Same as classifier_forward(filled=False) but keeps the autograd graph (params are leaf tensors).  bn=True: the
    model is in train() mode, i.e. BatchNorm uses minibatch statistics (`running` receives the updated buffers)."""
    masks = iter(relu_masks) if relu_masks is not None else None
    dmasks = iter(dropout_masks) if dropout_masks is not None else None
    if arch in ('resnet8', 'resnet16'):
        z = resnet_features(params, x.float(), arch, units, filled=False, bn=bn, bn_train=bn, running=running, relu_masks=masks,
                            dropout=dropout, dropout_masks=dmasks)
    else:
        sizes = {'conv31': [7, 5, 5], 'conv63': [7, 5, 5, 5], 'conv127': [7, 5, 5, 5, 5]}[arch]
        z = basicconv_features(params, x.float(), sizes, units, unit_scaling, filled=False, bn=bn, bn_train=bn,
                               running=running, act_masks=masks, dropout=dropout, dropout_masks=dmasks)
    return _conv(z, params['classifier.weight'], params['classifier.bias'])


# --------------------------------------------------------------------------------------
# Patch tiling for scoring, NMS
# --------------------------------------------------------------------------------------

def score_in_patches(fwd, X: torch.Tensor, patch_size: int, pad: int) -> np.ndarray:
    """predict_in_patches / get_patches / reconstruct_from_patches (model/utils.py:110-193), 2-D.
    ``fwd`` maps a [1,1,h,w] tensor to a [1,1,h,w] tensor. Returns float64 like the reference."""
    y, x = X.shape[-2:]
    Xp = F.pad(X, (pad, pad, pad, pad))
    yp, xp = Xp.shape[-2:]
    step = patch_size - 2 * pad
    out = np.zeros(tuple(X.shape))
    for i in range(0, y, step):
        for j in range(0, x, step):
            patch = Xp[..., i:min(i + patch_size, yp), j:min(j + patch_size, xp)]
            s = fwd(patch)[0, 0].numpy()[pad:-pad, pad:-pad]
            out[..., i:i + s.shape[-2], j:j + s.shape[-1]] = s
    return out


def nms(x: np.ndarray, r: int, threshold: float = -np.inf):
    """algorithms.non_maximum_suppression (algorithms.py:25-63), including the clip-to-shape quirk."""
    ii, jj = np.meshgrid(np.arange(-r, r + 1), np.arange(-r, r + 1))
    mask = (ii ** 2 + jj ** 2) <= r * r
    ii, jj = ii[mask], jj[mask]
    W = x.shape[1]
    A = x.ravel()
    I = np.argsort(A, axis=None)[::-1]
    S = np.zeros(len(A) + (x.shape[0] + 1) * W + W + 1, dtype=bool)
    scores, coords = [], []
    for i in I:
        if A[i] <= threshold:
            break
        if not S[i]:
            xx, yy = i % W, i // W
            scores.append(A[i])
            coords.append((xx, yy))
            yc = np.clip(yy + ii, 0, x.shape[0])
            xc = np.clip(xx + jj, 0, x.shape[1])
            S[yc * W + xc] = True
    return np.asarray(scores, dtype=np.float32), np.asarray(coords, dtype=np.int32).reshape(-1, 2)


def nms3d(x: np.ndarray, r, scale: float = 1.0, threshold: float = -np.inf):
    """algorithms.non_maximum_suppression_3d (algorithms.py:66-103): flat-index deltas over the ball of radius scale*r, no
    bounds handling (the reference's set of suppressed flat indices may hold out-of-range values; they never match)."""
    r = scale * r
    width = int(np.ceil(r))
    ax = np.arange(-width, width + 1)
    ii, jj, kk = np.meshgrid(ax, ax, ax)
    mask = (ii ** 2 + jj ** 2 + kk ** 2) <= r * r
    deltas = ii[mask] * (x.shape[1] * x.shape[2]) + jj[mask] * x.shape[2] + kk[mask]
    A = x.ravel()
    n = len(A)
    I = np.argsort(A, axis=None)[::-1]
    S = np.zeros(n, dtype=bool)
    scores, coords = [], []
    for i in I:
        if A[i] <= threshold:
            break
        if not S[i]:
            zz, yy, xx = np.unravel_index(i, x.shape)
            scores.append(A[i])
            coords.append((xx, yy, zz))
            t = i + deltas
            S[t[(t >= 0) & (t < n)]] = True
    return np.asarray(scores, dtype=np.float32), np.asarray(coords, dtype=np.int32).reshape(-1, 3)


# --------------------------------------------------------------------------------------
# Preprocessing: Fourier-crop downsample, GMM normalisation
# --------------------------------------------------------------------------------------

def downsample(x: np.ndarray, factor=1, shape=None) -> np.ndarray:
    """utils/image.py:38-61: rfft2, keep rows [0:m//2] + [-m//2:] and columns [0:n//2+1], rescale, irfft2."""
    if shape is None:
        shape = (int(x.shape[-2] / factor), int(x.shape[-1] / factor))
    m, n = shape
    F = np.fft.rfft2(x)
    F = np.concatenate([F[..., 0:m // 2, 0:n // 2 + 1], F[..., -m // 2:, 0:n // 2 + 1]], axis=0)
    F = F * ((n * m) / (x.shape[-2] * x.shape[-1]))
    return np.fft.irfft2(F, s=shape).astype(x.dtype)


def _beta_logpdf(p, a, b):
    import math
    xlogy = lambda c, v: 0.0 if c == 0 else c * math.log(v)
    return xlogy(a - 1, p) + xlogy(b - 1, 1 - p) + math.lgamma(a + b) - math.lgamma(a) - math.lgamma(b)


def gmm_fit(x: torch.Tensor, pi, split, alpha, beta, scale=1, tol=1e-3, num_iters=100):
    """stats.py:122-214 with share_var=True, in float32 tensors like the reference."""
    mu = torch.mean(x)
    pi = torch.as_tensor(pi)
    p0 = (x <= split).float()
    p1 = 1 - p0

    def m_step(p0, p1):
        s0, s1 = torch.sum(p0), torch.sum(p1)
        mu0 = torch.sum(x * p0) / s0 if s0 > 0 else mu
        mu1 = torch.sum(x * p1) / s1 if s1 > 0 else mu
        var = torch.mean(p0 * (x - mu0) ** 2 + p1 * (x - mu1) ** 2)
        return mu0, mu1, var

    def e_step(mu0, mu1, var, pi):
        l0 = -(x - mu0) ** 2 / 2 / var - 0.5 * torch.log(2 * np.pi * var) + torch.log1p(-pi)
        l1 = -(x - mu1) ** 2 / 2 / var - 0.5 * torch.log(2 * np.pi * var) + torch.log(pi)
        ma = torch.max(l0, l1)
        Z = ma + torch.log(torch.exp(l0 - ma) + torch.exp(l1 - ma))
        return l0, l1, Z, scale * torch.sum(Z) + _beta_logpdf(float(pi), alpha, beta)

    mu0, mu1, var = m_step(p0, p1)
    l0, l1, Z, logp = e_step(mu0, mu1, var, pi)
    logp_cur = logp
    for _ in range(1, num_iters + 1):
        p0, p1 = torch.exp(l0 - Z), torch.exp(l1 - Z)
        s = torch.sum(p1)
        a, b = alpha + s, beta + p1.numel() - s
        pi = (a - 1) / (a + b - 2)
        mu0, mu1, var = m_step(p0, p1)
        l0, l1, Z, logp = e_step(mu0, mu1, var, pi)
        if logp - logp_cur <= tol:
            break
        logp_cur = logp
    return logp, mu0, var, mu1, var, pi


def norm_fit(x: np.ndarray, alpha=900, beta=1, scale=1, num_iters=100):
    """stats.py:86-119."""
    pis = np.array([0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 0.95, 0.98, 1])
    splits = np.quantile(x, 1 - pis)
    logps, mus, stds = np.zeros(len(pis)), np.zeros(len(pis)), np.zeros(len(pis))
    xt = torch.from_numpy(np.ascontiguousarray(x))
    for i in range(len(pis)):
        if pis[i] == 1:
            mu, var = xt.mean(), xt.var()
            quirk = float(alpha) if beta == 1 else (0.0 if beta > 1 else np.inf)          # beta.PDF (not logpdf) at 1, stats.py:106
            logp = scale * torch.sum(-(xt - mu) ** 2 / 2 / var - 0.5 * torch.log(2 * np.pi * var)) + quirk
            pi = torch.as_tensor(1.0)
        else:
            logp, _, _, mu, var, pi = gmm_fit(xt, pis[i], splits[i], alpha, beta, scale, 1e-3, num_iters)
        pis[i], logps[i], mus[i], stds[i] = pi.item(), logp.item(), mu.item(), np.sqrt(var.item())
    i = int(np.argmax(logps))
    return mus[i], stds[i], pis[i], logps[i], mus, stds, pis, logps


def gmm_normalize(x: np.ndarray, alpha=900, beta=1, num_iters=100):
    """stats.normalize(method='gmm', sample=1) (stats.py:49-83) -> (normalised float32 image, mu, std, pi)."""
    mu, std, pi, *_ = norm_fit(x, alpha, beta, 1, num_iters)
    return ((x - mu) / std).astype(np.float32), mu, std, pi
