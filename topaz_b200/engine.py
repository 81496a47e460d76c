"""Execution engine: turns the drop-in nn.Modules into sequences of sm_100a kernel launches.

Dense ("filled") classifier forward and the U-Net denoiser forward run in fp16 operands / fp32 accumulation
on the tcgen05 implicit-GEMM kernel (same 11-bit significand as the TF32 cuDNN path the reference uses on
GPU).  The strided (training) classifier forward/backward lives in ``topaz_b200.train_engine``.

Plans (packed weights, k-block tables) are cached per module and invalidated when any parameter/buffer
changes (data_ptr or in-place version bump), or when fill()/unfill()/train()/eval() changes the geometry.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

import os

from . import ops
from .ops import ConvPart

# Engine knobs (environment, so the reference CLI signatures stay unchanged):
#   TPZ_FIRST=simt        Cin=1 first convs on the fp32 SIMT kernel instead of the tensor cores
#   TPZ_FIRST=im2col      Cin=1 first convs as HBM im2col + 1-tap tensor-core GEMM (the pre-fusion path)
#   TPZ_RESIDUAL=epilogue identity skip added in the epilogue (global loads) instead of an identity k-block in the MMA
FIRST_ON_TC = os.environ.get('TPZ_FIRST', 'tc') != 'simt'
FIRST_FUSED = os.environ.get('TPZ_FIRST', 'tc') == 'tc'      # im2col tile built in smem inside the GEMM kernel
RESIDUAL_IN_MMA = os.environ.get('TPZ_RESIDUAL', 'mma') != 'epilogue'
UP2_FUSED = os.environ.get('TPZ_UP2', 'fused') != 'off'       # fused 2x up-sampling in U-Net dec1.0 (poly-phase)
# Cout=1 U-Net tail: 'tc' = 16-column tensor-core GEMM with a fused dot epilogue, 'simt' = fp32 CUDA-core kernel;
# default: the tiled CUDA-core kernel where it exists (2-D, 32 channels, 3x3 / 5x5: 1600 FLOP/px is too little for the
# tensor-core path), the GEMM elsewhere (3-D)
LAST_MODE = os.environ.get('TPZ_LAST', 'auto')
# Strict inference (TPZ_PRECISION=strict, or engine.PRECISION = 'strict' at run time): every fp16 activation and weight is carried as a
# (hi, lo) pair -- 22 significand bits -- and each product as hi*hi + hi*lo + lo*hi on the same tcgen05 kernels (3x the MMAs,
# 2x the activation bytes).  It exists to SHOW that the kernels meet the north-star 1e-3 on every network, incl. the
# He-random seeded ones where the default 11-bit operands (= the TF32 of the reference's own GPU path) land at 1-3e-3.
#   TPZ_PRECISION = auto (default) | fast | strict   (engine.PRECISION at run time; TPZ_STRICT=1 is an alias of strict)
#     fast   : 11-bit operands everywhere (what the reference's own GPU path computes with its TF32 convolutions)
#     strict : split operands everywhere
#     auto   : fast, except where a PRETRAINED network of the reference was measured above 1e-3 in fast mode: the 3-D U-Net
#              (unet-3d-10a on N(0,1) input: 5e-3 -- its output is a 100x cancellation of O(1) features).  There the last
#              four convolutions (dec2.2, dec1.0, dec1.2, dec1.4), which carry 90 % of that error, run with split operands
#              (tools/precision_probe.py reproduces the per-layer error budget on the CPU).
PRECISION = os.environ.get('TPZ_PRECISION', 'strict' if os.environ.get('TPZ_STRICT', '0') == '1' else 'auto')
# fp16 range guard (default on): activations are stored multiplied by a power of two chosen from max|input| (ops.range_scale)
RANGE_GUARD = os.environ.get('TPZ_RANGE_GUARD', '1') != '0'
# Dense classifier forward through the model-level C ABI (csrc/tpz_model.cu: plans + on-device weight repack in C++);
# 'py' keeps the Python-built plans (same kernels, same packed bytes).  Strict precision always uses the Python plans.
DENSE_ENGINE = os.environ.get('TPZ_DENSE_ENGINE', 'c')
# U-Net denoiser forward through the model-level C ABI (csrc/tpz_unet.cu: tpz_unet2d_forward / tpz_unet3d_forward): opt-in with
# TPZ_UNET_ENGINE=c.  Same kernels, same packed bytes and the same launch sequence as the Python plans (tests/test_unet_abi.py
# holds them bit-identical on the CPU simulation of the kernels); 'py' stays the default until the C path has run on hardware.
UNET_ENGINE = os.environ.get('TPZ_UNET_ENGINE', 'py')


def _rup(c: int, m: int = 32) -> int:
    return (c + m - 1) // m * m


def _tap_ld(taps: int) -> int:
    """channel count of an im2col tensor: next power of two >= 32 (the im2col kernel indexes chunks with shifts)."""
    ld = 32
    while ld < taps:
        ld *= 2
    return ld


def _slope_of(act) -> float:
    if isinstance(act, nn.ReLU):
        return 0.0
    if isinstance(act, nn.LeakyReLU):
        return float(act.negative_slope)
    if isinstance(act, nn.PReLU):
        if act.weight.numel() != 1:
            raise NotImplementedError('topaz_b200: per-channel PReLU is not supported')
        return float(act.weight.detach().reshape(-1)[0])
    if isinstance(act, nn.Identity):
        return 1.0
    raise NotImplementedError(f'topaz_b200: unsupported activation {type(act).__name__}')


def _bn_affine(bn: Optional[nn.Module], co: int):
    """eval-mode BatchNorm as y = a*x + b (a = gamma/sqrt(var+eps), b = beta - a*mean)."""
    if bn is None:
        return None, None
    a = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    b = bn.bias.detach().float() - a * bn.running_mean.detach().float()
    return a.cpu(), b.cpu()


def _state_key(module: nn.Module, extra=()):
    ks = [(p.data_ptr(), p._version) for p in module.parameters()]
    ks += [(b.data_ptr(), b._version) for b in module.buffers()]
    return (tuple(ks), module.training, module.__dict__.get('_tpz_epoch', 0), PRECISION) + tuple(extra)


def _cached(module: nn.Module, name: str, key, build):
    cache = module.__dict__.setdefault('_tpz_plans', {})
    hit = cache.get(name)
    if hit is not None and hit[0] == key:
        return hit[1]
    plan = build()
    cache[name] = (key, plan)
    return plan


# ------------------------------------------------------------------------------------------------
# classifier (ResNet8/16, conv31/63/127) dense forward
# ------------------------------------------------------------------------------------------------
def _feature_blocks(features, slopes: bool = True, dropout: bool = False) -> List[dict]:
    """Describe the feature extractor's layers in their CURRENT geometry (after fill()/unfill()).  slopes=False (training)
    leaves the activation slopes unread -- a PReLU slope lives on the device and reading it would synchronise every step;
    the activation modules themselves are returned under 'act' / 'act0' / 'act1'.  dropout=True (training engine) lists the
    active nn.Dropout layers as {'kind': 'dropout', 'p': p} blocks; otherwise an active dropout layer is an error (the dense
    path runs in eval() mode, where dropout is the identity)."""
    _slope = _slope_of if slopes else (lambda act: None)
    from .model.features import resnet as R, basic as B
    blocks = []
    if isinstance(features, R.ResNet):
        for mod in features.features.children():
            if isinstance(mod, R.BasicConv):
                blocks.append(dict(kind='conv', w=mod.conv.weight, b=mod.conv.bias, bn=getattr(mod, 'bn', None),
                                   dil=mod.conv.dilation[0], stride=mod.conv.stride[0], slope=_slope(mod.act), act=mod.act))
            elif isinstance(mod, R.ResidA):
                blocks.append(dict(kind='resid', w0=mod.conv0.weight, b0=mod.conv0.bias, bn0=getattr(mod, 'bn0', None),
                                   w1=mod.conv1.weight, b1=mod.conv1.bias, bn1=getattr(mod, 'bn1', None),
                                   proj=mod.proj.weight if hasattr(mod, 'proj') else None,
                                   d0=mod.conv0.dilation[0], d1=mod.conv1.dilation[0], stride=mod.conv1.stride[0],
                                   slope0=_slope(mod.act0), slope1=_slope(mod.act1), act0=mod.act0, act1=mod.act1))
            elif isinstance(mod, nn.Dropout):
                if mod.training and mod.p > 0:
                    if not dropout:
                        raise NotImplementedError('topaz_b200: the dense forward expects eval() mode (active dropout layer)')
                    blocks.append(dict(kind='dropout', p=float(mod.p)))
            else:
                raise NotImplementedError(f'topaz_b200: unsupported ResNet child {type(mod).__name__}')
    elif isinstance(features, B.BasicConv):
        mods = list(features.features.children())
        i = 0
        while i < len(mods):
            conv = mods[i]; i += 1
            if isinstance(conv, nn.Dropout):
                if conv.training and conv.p > 0:
                    if not dropout:
                        raise NotImplementedError('topaz_b200: the dense forward expects eval() mode (active dropout layer)')
                    blocks.append(dict(kind='dropout', p=float(conv.p)))
                continue
            assert isinstance(conv, (nn.Conv2d, nn.Conv3d)), type(conv)
            bn = None
            if i < len(mods) and isinstance(mods[i], (nn.BatchNorm2d, nn.BatchNorm3d)):
                bn = mods[i]; i += 1
            act = mods[i]; i += 1
            blocks.append(dict(kind='conv', w=conv.weight, b=conv.bias, bn=bn, dil=conv.dilation[0],
                               stride=conv.stride[0], slope=_slope(act), act=act))
    else:
        raise NotImplementedError(f'topaz_b200: unsupported feature extractor {type(features).__name__}')
    return blocks


def is_filled(features) -> bool:
    return bool(getattr(features, 'pad', False) or getattr(features, 'filled', False))


def _build_dense_plan(features, classifier: Optional[nn.Module], device):
    if getattr(features, 'dims', 2) != 2:
        raise NotImplementedError('topaz_b200: 3-D classifiers are outside the B200 hot path')
    blocks = _feature_blocks(features)
    if features.training and any(b.get('bn') is not None or b.get('bn0') is not None for b in blocks):
        raise NotImplementedError('topaz_b200: dense forward with BatchNorm requires eval() mode')
    steps = []
    first = blocks[0]
    assert first['kind'] == 'conv' and first['w'].shape[1] == 1, 'first layer must be a Cin=1 conv'
    if any(b['stride'] != 1 for b in blocks):
        raise RuntimeError('topaz_b200: dense plan requested on an unfilled (strided) model')
    a, sh = _bn_affine(first['bn'], first['w'].shape[0])
    w = first['w'].detach().float().cpu()
    b = first['b'].detach().float().cpu() if first['b'] is not None else torch.zeros(w.shape[0])
    if a is not None:
        w = w * a.view(-1, 1, 1, 1); b = b * a + sh
    c_real = w.shape[0]
    k0 = w.shape[-1]
    strict = PRECISION == 'strict'
    pack = lambda *a, **k: ops.pack_tc_conv(*a, strict=strict, **k)
    _CP = ConvPart
    ConvPart_ = lambda *a, **k: _CP(*a, split=strict, **k)
    if not strict and FIRST_FUSED and first['dil'] == 1 and ops.first_tc_supported(k0, _rup(c_real)):
        # Cin = 1: one tcgen05 kernel that builds the im2col tile in shared memory (tpz_first_tc.cu)
        wp, bp = ops.pack_first_tc(w[:, 0], b, _rup(c_real), device)
        steps.append(dict(op='first_tc', w=wp, b=bp, k=k0, pad=features.width // 2, slope=first['slope']))
    elif not strict and FIRST_ON_TC and first['dil'] == 1 and k0 * k0 <= 128:
        # Cin = 1: im2col (taps -> channels) then a 1-tap tensor-core GEMM (K = k*k padded to 32)
        ld = _tap_ld(k0 * k0)
        p1 = pack([ConvPart(w.reshape(c_real, k0 * k0, 1, 1), ld, 1)], b, _rup(c_real), first['slope'], device)
        steps.append(dict(op='im2col', k=k0, pad=features.width // 2, ld=ld))
        steps.append(dict(op='tc', plan=p1, shrink=0, src='cur', dot=False, save_in=False))
    else:
        steps.append(dict(op='first', w=w[:, 0][:, None].contiguous().to(device), b=b.to(device), dil=first['dil'],
                          pad=features.width // 2, slope=first['slope'], out_ld=_rup(c_real)))
    c_store = _rup(c_real)
    nblk = len(blocks)
    for bi, blk in enumerate(blocks[1:], 1):
        last = (bi == nblk - 1)
        if blk['kind'] == 'conv':
            w = blk['w']; co = w.shape[0]
            a, sh = _bn_affine(blk['bn'], co)
            bias = blk['b'].detach().float().cpu() if blk['b'] is not None else torch.zeros(co)
            if a is not None:
                bias = bias * a + sh
            fuse = last and classifier is not None
            plan = pack([ConvPart_(w, c_store, blk['dil'])], bias, _rup(co), blk['slope'], device,
                                    out_scale=a,
                                    dot_w=classifier.weight if fuse else None,
                                    dot_b=float(classifier.bias.detach()[0]) if fuse else 0.0)
            k = w.shape[-1]
            steps.append(dict(op='tc', plan=plan, shrink=(k - 1) * blk['dil'], src='cur', dot=fuse, save_in=False))
            c_store = _rup(co)
        else:
            w0, w1 = blk['w0'], blk['w1']
            ch, co = w0.shape[0], w1.shape[0]
            a0, s0 = _bn_affine(blk['bn0'], ch)
            b0 = blk['b0'].detach().float().cpu() if blk['b0'] is not None else torch.zeros(ch)
            if a0 is not None:
                b0 = b0 * a0 + s0
            p0 = pack([ConvPart_(w0, c_store, blk['d0'])], b0, _rup(ch), blk['slope0'], device, out_scale=a0)
            steps.append(dict(op='tc', plan=p0, shrink=2 * blk['d0'], src='cur', dot=False, save_in=True))
            a1, s1 = _bn_affine(blk['bn1'], co)
            b1 = blk['b1'].detach().float().cpu() if blk['b1'] is not None else torch.zeros(co)
            if a1 is not None:
                b1 = b1 * a1 + s1
            edge = blk['d0'] + blk['d1']
            parts = [ConvPart_(w1, _rup(ch), blk['d1'])]
            if blk['proj'] is not None:
                parts.append(ConvPart_(blk['proj'], c_store, 1, (edge, edge, 0)))
                p1 = pack(parts, b1, _rup(co), blk['slope1'], device, out_scale=a1)
                steps.append(dict(op='tc', plan=p1, shrink=2 * blk['d1'], src='cur+saved', dot=False, save_in=False))
            elif RESIDUAL_IN_MMA or strict:
                # identity skip as one extra k-block per channel chunk: A = cropped block input, B = I (exact: x*1.0
                # accumulated in fp32), so the epilogue issues no global loads
                eye = torch.eye(co, dtype=torch.float32).reshape(co, co, 1, 1)
                p1 = pack(parts, b1, _rup(co), blk['slope1'], device, out_scale=a1)
                pe = pack([ConvPart_(w1, _rup(ch), blk['d1']), ConvPart_(eye, c_store, 1, (edge, edge, 0))], b1,
                                      _rup(co), blk['slope1'], device, out_scale=a1)
                steps.append(dict(op='tc', plan=pe, shrink=2 * blk['d1'], src='cur+saved', dot=False, save_in=False))
            else:
                p1 = pack(parts, b1, _rup(co), blk['slope1'], device, out_scale=a1, res_scale=a1)
                steps.append(dict(op='tc', plan=p1, shrink=2 * blk['d1'], src='cur', res=True, edge=edge, dot=False,
                                  save_in=False))
            c_store = _rup(co)
    return dict(steps=steps, strict=strict,
                c_out=blocks[-1]['w'].shape[0] if blocks[-1]['kind'] == 'conv' else blocks[-1]['w1'].shape[0])


def _run_dense(plan, x: torch.Tensor, want_features: bool):
    """x: fp32 [B, H, W] contiguous on device.  Returns (fp32 [B,1,H,W] logits (dot fused) or fp16 NDHWC features still
    multiplied by rng[0], rng)."""
    B, H, W = x.shape
    cur = None
    saved = None
    out = None
    strict = plan['strict']
    rng = ops.range_scale(x) if RANGE_GUARD else None
    for st in plan['steps']:
        if st['op'] == 'first':
            cur = ops.conv_first(x.view(B, 1, H, W), st['w'], st['b'], st['dil'], st['pad'], st['slope'], st['out_ld'],
                                 rng=rng, split=strict)
            continue
        if st['op'] == 'first_tc':
            cur = ops.conv_first_tc(x, st['w'], st['b'], st['k'], st['pad'], st['slope'], rng=rng)
            continue
        if st['op'] == 'im2col':
            cur = ops.im2col_first(x, st['k'], st['pad'], st['ld'], rng=rng)
            continue
        p = st['plan']
        N, D, Hc, Wc, _ = cur.shape
        Ho, Wo = Hc - st['shrink'], Wc - st['shrink']
        if st['save_in']:
            saved = cur
        srcs = [cur, saved] if st['src'] == 'cur+saved' else [cur]
        res = saved if st.get('res') else None
        res_org = (st['edge'], st['edge'], 0) if st.get('res') else (0, 0, 0)
        if st['dot'] and not want_features:
            out = torch.empty((N, D, Ho, Wo), dtype=torch.float32, device=x.device)
            ops.tc_conv(p, srcs, (N, D, Ho, Wo), out=None, res=res, res_org=res_org, dot_out=out, tag='dominant', rng=rng)
            cur = None
        else:
            nxt = torch.empty((N, D, Ho, Wo, p.out_channels), dtype=torch.float16, device=x.device)
            ops.tc_conv(p, srcs, (N, D, Ho, Wo), out=nxt, res=res, res_org=res_org, rng=rng)
            cur = nxt
    return (out if out is not None else cur), rng


def _as_image_batch(x: torch.Tensor, dims: int) -> torch.Tensor:
    ops.require_cuda(x, 'classifier input')
    if x.dim() < dims + 2:
        x = x.unsqueeze(1)
    if x.shape[1] != 1:
        raise ValueError('topaz_b200: expected a single input channel')
    return x[:, 0].contiguous().float()


def classifier_forward(model, x: torch.Tensor) -> torch.Tensor:
    """LinearClassifier.forward (reference classifier.py:48-66)."""
    feats = model.features
    if not is_filled(feats):
        from . import train_engine
        return train_engine.classifier_forward(model, x)
    xi = _as_image_batch(x, 2)
    key = _state_key(model, ('dense', str(xi.device)))
    if DENSE_ENGINE == 'c' and PRECISION != 'strict' and RANGE_GUARD and ops.TC_VARIANT == 'auto':
        dm = _dense_c_model(model, key)
        if dm is not None:               # None: a weight row needs the row-scaled plans, which only the Python packer builds
            return dm.forward(xi)
    plan = _cached(model, 'dense_cls', key, lambda: _build_dense_plan(feats, model.classifier, xi.device))
    y, _ = _run_dense(plan, xi, want_features=False)   # [B, 1(D), H, W]; the fused dot epilogue already undid the range scale
    return y.view(xi.shape[0], 1, y.shape[2], y.shape[3])


def _dense_c_model(model, key):
    """The model's native handle (model_abi.DenseModel), repacked on the device when a parameter changed and rebuilt when the
    geometry did (fill()/unfill() change dilations)."""
    from .model_abi import DenseModel, WeightRangeError
    cache = model.__dict__.setdefault('_tpz_plans', {})
    hit = cache.get('dense_c')
    geom = tuple((b.get('dil'), b.get('d0'), b.get('d1')) for b in _feature_blocks(model.features, slopes=False))
    if hit is not None and hit[0] == key:
        return hit[2]
    try:
        if hit is not None and hit[1] == geom and hit[2] is not None:
            hit[2].update()
            cache['dense_c'] = (key, geom, hit[2])
            return hit[2]
        if hit is not None and hit[2] is not None:
            hit[2].close()
        dm = DenseModel(model)
    except WeightRangeError:
        if hit is not None and hit[2] is not None:
            hit[2].close()
        dm = None                        # remembered for this parameter state: the Python plans (ops._row_scales) take over
    cache['dense_c'] = (key, geom, dm)
    return dm


def features_forward(features, x: torch.Tensor) -> torch.Tensor:
    """ResNet.forward / BasicConv.forward without the classifier head: NCHW fp32 feature map."""
    if not is_filled(features):
        from . import train_engine
        return train_engine.features_forward(features, x)
    xi = _as_image_batch(x, 2)
    key = _state_key(features, ('dense_feat', str(xi.device)))
    plan = _cached(features, 'dense_feat', key, lambda: _build_dense_plan(features, None, xi.device))
    z, rng = _run_dense(plan, xi, want_features=True)   # [B,1,H,W,Cstore] fp16 (strict: [hi | lo] halves), scaled by rng[0]
    c = plan['c_out']
    f = z[:, 0, :, :, :c].float()
    if plan['strict']:
        half = z.shape[-1] // 2
        f = f + z[:, 0, :, :, half:half + c].float()
    if rng is not None:
        f = f * rng[1]
    return f.permute(0, 3, 1, 2).contiguous()


# ------------------------------------------------------------------------------------------------
# U-Net denoisers (UDenoiseNet / UDenoiseNet3D / UDenoiseNetSmall)
# ------------------------------------------------------------------------------------------------
def _conv_of(seq, idx):
    c = seq[idx]
    assert isinstance(c, (nn.Conv2d, nn.Conv3d))
    return c


def _up2_phase_plans(w_up, up_c, k, second_part, bias, co_store, slope, device, dims=2, strict=False, up_split=False,
                     split_out=False):
    """Plans for conv(cat[nearest_up2(h), other]) computed per output phase directly from the half-resolution tensor
    h: taps of the k^dims kernel that alias onto the same half-res voxel are summed (5x5 -> 3x3, 3x3(x3) -> 2x2(x2) per
    phase).  ``second_part(phase)`` returns the ConvPart of the full-resolution source for phase = (px, py, pz)."""
    pad = k // 2
    wu = w_up.detach().float().cpu()
    if wu.dim() == 4:
        wu = wu[:, :, None]
    kz = wu.shape[2]
    plans = []
    for pz in range(2 if dims == 3 else 1):
        offz = [(pz + r - pad) // 2 for r in range(kz)] if dims == 3 else [0]
        for py in range(2):
            offy = [(py + r - pad) // 2 for r in range(k)]
            for px in range(2):
                offx = [(px + r - pad) // 2 for r in range(k)]
                az0, ay0, ax0 = min(offz), min(offy), min(offx)
                wm = torch.zeros((wu.shape[0], up_c, max(offz) - az0 + 1, max(offy) - ay0 + 1, max(offx) - ax0 + 1))
                for q in range(kz):
                    for r in range(k):
                        for t in range(k):
                            wm[:, :, offz[q] - az0, offy[r] - ay0, offx[t] - ax0] += wu[:, :, q, r, t]
                parts = [ConvPart(wm, _rup(up_c), 1, (ax0, ay0, az0), lat=1, lat_z=1, phase=False, split=up_split),
                         second_part((px, py, pz))]
                plans.append(ops.pack_tc_conv(parts, bias, co_store, slope, device, lattice=2, phase_sel=py * 2 + px + 1,
                                              lattice_z=2 if dims == 3 else 1, phase_z=pz, strict=strict, split_out=split_out))
    return plans


def _unet_precision(dims: int, depth: int):
    """Which U-Net layers run with split (hi, lo) operands, and which tensors must therefore be stored as (hi, lo) pairs.
    Layers are named ('enc', i, 0) / ('dec', l, 0|2|4) after the reference's Sequential indices; the raw input is 'raw'."""
    ndec = depth - 1
    layers = [('enc', i, 0) for i in range(1, depth + 1)]
    for l in range(ndec, 0, -1):
        layers += [('dec', l, 0), ('dec', l, 2)]
    layers.append(('dec', 1, 4))
    if PRECISION == 'strict':
        S = set(layers)
    elif PRECISION == 'auto' and dims == 3 and ndec >= 2:
        S = {('dec', 2, 2), ('dec', 1, 0), ('dec', 1, 2), ('dec', 1, 4)}
    else:
        S = set()
    cons = {}
    for i in range(1, depth):
        cons.setdefault(('enc', i, 0), []).append(('enc', i + 1, 0))
    for l in range(ndec, 0, -1):
        cons.setdefault(('enc', depth, 0) if l == ndec else ('dec', l + 1, 2), []).append(('dec', l, 0))
        cons.setdefault(('enc', l - 1, 0) if l > 1 else 'raw', []).append(('dec', l, 0))
        cons.setdefault(('dec', l, 0), []).append(('dec', l, 2))
    cons.setdefault(('dec', 1, 2), []).append(('dec', 1, 4))
    split = {t: any(c in S for c in cs) for t, cs in cons.items()}
    return S, split


def _build_unet_plan(model, device):
    enc = [getattr(model, f'enc{i}') for i in range(1, 10) if hasattr(model, f'enc{i}')]
    depth = len(enc)
    ndec = depth - 1
    dec = {l: getattr(model, f'dec{l}') for l in range(ndec, 0, -1)}
    dims = 3 if isinstance(enc[0][0], nn.Conv3d) else 2
    nf = enc[0][0].weight.shape[0]
    slope = 0.1

    def same_org(k):
        p = k // 2
        return (-p, -p, -p if dims == 3 else 0)

    S, split = _unet_precision(dims, depth)

    def pk(layer, parts, *a, **k):
        return ops.pack_tc_conv(parts, *a, strict=layer in S, split_out=split.get(layer, False), **k)

    plan = dict(dims=dims, nf=nf, nenc=depth, split=split, strict_layers=S)
    c1 = enc[0][0]
    k1 = c1.weight.shape[-1]
    plan['first_tc'] = None
    plan['first_fused'] = None
    e1 = ('enc', 1, 0)
    if e1 in S or split.get(e1, False):
        pass                                              # Cin = 1 first conv on the fp32 CUDA-core kernel, (hi, lo) output
    elif FIRST_FUSED and dims == 2 and ops.first_tc_supported(k1, _rup(nf)):
        wp, bp = ops.pack_first_tc(c1.weight.detach()[:, 0], c1.bias, _rup(nf), device)
        plan['first_fused'] = dict(w=wp, b=bp, k=k1)
    elif FIRST_ON_TC and k1 * k1 <= 128:
        # Cin = 1: in-plane im2col (k*k taps -> channels) + tensor-core GEMM; in 3-D the k z-taps stay taps of the GEMM
        ld = _tap_ld(k1 * k1)
        if dims == 2:
            w1 = c1.weight.detach().reshape(nf, k1 * k1, 1, 1)
            org = (0, 0, 0)
        else:
            w1 = c1.weight.detach().reshape(nf, k1, k1 * k1).permute(0, 2, 1).reshape(nf, k1 * k1, k1, 1, 1)
            org = (0, 0, -(k1 // 2))
        plan['first_tc'] = dict(k=k1, ld=ld, plan=ops.pack_tc_conv([ConvPart(w1, ld, 1, org)], c1.bias, _rup(nf), slope, device))
    plan['first'] = dict(w=c1.weight.detach().float().reshape((nf,) + ((1,) if dims == 2 else ()) + tuple(c1.weight.shape[2:])).contiguous().to(device),
                         b=c1.bias.detach().float().to(device), pad=c1.weight.shape[-1] // 2, out_ld=_rup(nf),
                         pool=len(enc[0]) > 2, split=split.get(e1, False))
    plan['enc'] = []
    for i, e in enumerate(enc[1:], 2):
        c = e[0]
        k = c.weight.shape[-1]
        name = ('enc', i, 0)
        p = pk(name, [ConvPart(c.weight, _rup(nf), 1, same_org(k), split=name in S)], c.bias, _rup(c.weight.shape[0]), slope, device)
        plan['enc'].append(dict(plan=p, pool=len(e) > 2))
    plan['dec'] = {}
    up_c = nf
    for l in range(ndec, 0, -1):
        d = dec[l]
        ca, cb = d[0], d[2]
        k = ca.weight.shape[-1]
        na, nb = ('dec', l, 0), ('dec', l, 2)
        sa, sb = na in S, nb in S
        if l > 1:
            skip_c = nf
            parts = [ConvPart(ca.weight[:, :up_c], _rup(up_c), 1, same_org(k), split=sa),
                     ConvPart(ca.weight[:, up_c:], _rup(skip_c), 1, same_org(k), split=sa)]
            pa = pk(na, parts, ca.bias, _rup(ca.weight.shape[0]), slope, device)
            pb = pk(nb, [ConvPart(cb.weight, _rup(ca.weight.shape[0]), 1, same_org(cb.weight.shape[-1]), split=sb)],
                    cb.bias, _rup(cb.weight.shape[0]), slope, device)
            up2 = None
            if UP2_FUSED:
                w_skip, pad_k = ca.weight[:, up_c:], k // 2
                up2 = _up2_phase_plans(ca.weight[:, :up_c], up_c, k,
                                       lambda ph: ConvPart(w_skip, _rup(skip_c), 1,
                                                           (ph[0] - pad_k, ph[1] - pad_k, ph[2] - pad_k if dims == 3 else 0),
                                                           lat=2, lat_z=2 if dims == 3 else 0, phase=False, split=sa),
                                       ca.bias, _rup(ca.weight.shape[0]), slope, device, dims, strict=sa, up_split=sa,
                                       split_out=split.get(na, False))
            plan['dec'][l] = dict(a=pa, b=pb, up2=up2)
            up_c = cb.weight.shape[0]
        else:
            # dec1: [upsampled (up_c ch), raw image (1 ch)] -> conv,lrelu,conv,lrelu,conv
            ntap = k ** dims
            wraw = ca.weight[:, up_c].reshape(ca.weight.shape[0], ntap)         # [Co, taps]
            raw_part = ConvPart(wraw.reshape(wraw.shape[0], ntap, 1, 1, 1), _tap_ld(ntap), 1, (0, 0, 0), split=sa)
            parts = [ConvPart(ca.weight[:, :up_c], _rup(up_c), 1, same_org(k), split=sa), raw_part]
            pa = pk(na, parts, ca.bias, _rup(ca.weight.shape[0]), slope, device)
            onehot = torch.eye(ntap, dtype=torch.float32).reshape((ntap,) + ((1,) if dims == 2 else ()) + (k,) * dims)
            pb = pk(nb, [ConvPart(cb.weight, _rup(ca.weight.shape[0]), 1, same_org(cb.weight.shape[-1]), split=sb)],
                    cb.bias, _rup(cb.weight.shape[0]), slope, device)
            cc = d[4]
            kl = cc.weight.shape[-1]
            cin = cc.weight.shape[1]
            wl = torch.zeros((kl ** dims, _rup(cin)), dtype=torch.float32)
            wl[:, :cin] = cc.weight.detach().float().cpu()[0].reshape(cin, -1).t()
            # fused nearest-2x up-sampling (2-D, exact factor 2): output phase (py,px) of the conv over the up-sampled
            # tensor equals a conv over the HALF-resolution tensor with the taps that alias onto the same source pixel
            # summed (5x5 -> 3x3 per phase: 2.8x fewer MACs on the dominant layer, and no up-sampled tensor in HBM).
            up2 = None
            if UP2_FUSED:
                up2 = _up2_phase_plans(ca.weight[:, :up_c], up_c, k,
                                       lambda ph: ConvPart(wraw.reshape(wraw.shape[0], ntap, 1, 1, 1), _tap_ld(ntap), 1,
                                                           (ph[0], ph[1], ph[2] if dims == 3 else 0), lat=2,
                                                           lat_z=2 if dims == 3 else 0, phase=False, split=sa),
                                       ca.bias, _rup(ca.weight.shape[0]), slope, device, dims, strict=sa, up_split=sa,
                                       split_out=split.get(na, False))
            # dec1.4 (Cout = 1) on the tensor-core kernel: 16 output columns (1 real), the fused "dot" epilogue picks
            # column 0, adds the bias and de-normalises -> dense fp32 image; no 16-channel tensor is written
            onehot0 = torch.zeros(1, 1, 1, 1); onehot0[0, 0, 0, 0] = 1.0
            nl = ('dec', 1, 4)
            pl = ops.pack_tc_conv([ConvPart(cc.weight, _rup(cin), 1, same_org(kl), split=nl in S)], None, 16, 1.0, device,
                                  dot_w=onehot0, dot_b=float(cc.bias.detach()[0]), strict=nl in S, split_out=False)
            plan['dec'][1] = dict(last_tc=pl, up2=up2, a=pa, b=pb, onehot=onehot.contiguous().to(device), k=k, ntap_store=_tap_ld(ntap),
                                  last_w=wl.contiguous().to(device), last_w2=torch.cat([wl, wl], dim=1).contiguous().to(device),
                                  last_b=float(cc.bias.detach()[0]),
                                  last_k=(kl if dims == 3 else 1, kl, kl), last_pad=kl // 2, last_c=cin,
                                  last_strict=nl in S, raw_split=split.get('raw', False))
    return plan


def _unet_c_model(model, key):
    """The denoiser's native handle (model_abi.UnetModel), or None where only the Python plans apply: a non-default kernel
    selection (the TPZ_FIRST / TPZ_UP2 / TPZ_LAST / TPZ_RANGE_GUARD switches), or a weight row that needs the row-scaled plans.
    One handle per (parameter state, precision); handles are kept for the life of the module -- a CUDA graph captured by
    Denoise._denoise_crop holds the addresses of a handle's packed weights and may be replayed whenever its key recurs
    (eval() / train() / eval(), a precision switched back), so a handle is never freed under a graph.  Denoiser parameters do not
    change on this path (denoiser training is out of scope), so the set stays small."""
    from .model_abi import UnetModel, WeightRangeError
    cache = model.__dict__.setdefault('_tpz_plans', {})
    defaults = bool(RANGE_GUARD and UP2_FUSED and FIRST_FUSED and LAST_MODE == 'auto' and ops.TC_VARIANT == 'auto')
    key = (key, defaults)
    hit = cache.get('unet_c')
    if hit is not None and hit[0] == key:
        return hit[1]
    handles = cache.setdefault('unet_c_handles', {})
    if key not in handles:
        um = None
        if defaults and PRECISION in UnetModel.PRECISIONS:
            try:
                um = UnetModel(model, precision=PRECISION)
            except (WeightRangeError, NotImplementedError):
                um = None
        handles[key] = um
    um = handles[key]
    cache['unet_c'] = (key, um)
    return um


def unet_forward(model, x: torch.Tensor, denorm_stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """UDenoiseNet(3D).forward (reference denoising/models.py:130-175, 508-564).
    x: fp32 [N,1,(D),H,W] on device.  If ``denorm_stats`` (device float[2]) is given the output is
    de-normalised (y*std+mean) inside the last kernel (denoise.py:295)."""
    ops.require_cuda(x, 'denoiser input')
    key = _state_key(model, ('unet', str(x.device)))
    if UNET_ENGINE == 'c':
        um = _unet_c_model(model, key)
        if um is not None:
            if x.dim() != um.dims + 2 or x.shape[1] != 1:
                raise ValueError(f'topaz_b200: expected input [N,1,{"D,H,W" if um.dims == 3 else "H,W"}], got {tuple(x.shape)}')
            return um.forward(x, denorm_stats)
    plan = _cached(model, 'unet', key, lambda: _build_unet_plan(model, x.device))
    dims = plan['dims']
    if x.dim() != dims + 2 or x.shape[1] != 1:
        raise ValueError(f'topaz_b200: expected input [N,1,{"D,H,W" if dims == 3 else "H,W"}], got {tuple(x.shape)}')
    xi = x[:, 0].contiguous().float()
    if dims == 2:
        xi = xi[:, None]                                  # [N, 1, H, W]
    f = plan['first']
    split = plan['split']
    rng = ops.range_scale(xi) if RANGE_GUARD else None
    fused_pool = False
    if plan['first_fused'] is not None:
        ff = plan['first_fused']
        h = ops.conv_first_tc(xi[:, 0].contiguous(), ff['w'], ff['b'], ff['k'], ff['k'] // 2, 0.1, pool=bool(f['pool']), rng=rng)
        fused_pool = bool(f['pool'])
    elif plan['first_tc'] is not None:
        ft = plan['first_tc']
        N0, D0, H0, W0 = xi.shape
        col = ops.im2col_first(xi.reshape(N0 * D0, H0, W0), ft['k'], ft['k'] // 2, ft['ld'], rng=rng).view(N0, D0, H0, W0, ft['ld'])
        h = torch.empty((N0, D0, H0, W0, ft['plan'].Co), dtype=torch.float16, device=x.device)
        ops.tc_conv(ft['plan'], [col], (N0, D0, H0, W0), out=h, rng=rng)
    else:
        h = ops.conv_first(xi, f['w'], f['b'], 1, f['pad'], 0.1, f['out_ld'], rng=rng, split=f['split'])
    skips = []
    if f['pool'] and not fused_pool:
        h = ops.maxpool2(h, dims, split=f['split'])
    skips.append(h)
    for e in plan['enc']:
        N, D, H, W, _ = h.shape
        o = torch.empty((N, D, H, W, e['plan'].out_channels), dtype=torch.float16, device=x.device)
        ops.tc_conv(e['plan'], [h], (N, D, H, W), out=o, rng=rng)
        h = ops.maxpool2(o, dims, split=e['plan'].split_out) if e['pool'] else o
        if e['pool']:
            skips.append(h)
    # skips = [p1, ..., p_{n-1}]; decoder level l joins p_{l-1} (level 1 joins the raw image)
    ndec = plan['nenc'] - 1
    for l in range(ndec, 0, -1):
        d = plan['dec'][l]
        if l > 1:
            skip = skips[l - 2]
            N, D, H, W, _ = skip.shape
            o = torch.empty((N, D, H, W, d['a'].out_channels), dtype=torch.float16, device=x.device)
            if d.get('up2') is not None and H == 2 * h.shape[2] and W == 2 * h.shape[3] and (dims == 2 or D == 2 * h.shape[1]):
                for pl2 in d['up2']:                      # fused nearest-2x up-sampling, one launch per output phase
                    ops.tc_conv(pl2, [h, skip], (N, D, H, W), out=o, rng=rng)
            else:
                up = ops.upsample_nearest(h, (D, H, W))
                ops.tc_conv(d['a'], [up, skip], (N, D, H, W), out=o, rng=rng)
            o2 = torch.empty((N, D, H, W, d['b'].out_channels), dtype=torch.float16, device=x.device)
            ops.tc_conv(d['b'], [o], (N, D, H, W), out=o2, rng=rng)
            h = o2
        else:
            N, D, H, W = xi.shape
            rs = d['raw_split']
            if dims == 2:
                raw = ops.im2col_first(xi[:, 0], d['k'], d['k'] // 2, d['ntap_store'], rng=rng, split=rs)
            else:
                raw = ops.im2col3d_first(xi, d['k'], d['ntap_store'], rng=rng, split=rs) if d['ntap_store'] % 16 == 0 else \
                    ops.conv_first(xi, d['onehot'], None, 1, d['k'] // 2, 1.0, d['ntap_store'], rng=rng, split=rs)
            o = torch.empty((N, D, H, W, d['a'].out_channels), dtype=torch.float16, device=x.device)
            if d.get('up2') is not None and H == 2 * h.shape[2] and W == 2 * h.shape[3] and (dims == 2 or D == 2 * h.shape[1]):
                for pl2 in d['up2']:                      # one launch per output phase, reading the half-res tensor
                    ops.tc_conv(pl2, [h, raw], (N, D, H, W), out=o, rng=rng)
            else:
                up = ops.upsample_nearest(h, (D, H, W))
                ops.tc_conv(d['a'], [up, raw], (N, D, H, W), out=o, rng=rng)
            o2 = torch.empty((N, D, H, W, d['b'].out_channels), dtype=torch.float16, device=x.device)
            ops.tc_conv(d['b'], [o], (N, D, H, W), out=o2, rng=rng)
            # Cout = 1 tail on the CUDA cores (fp32 math on the fp16 inputs): 2-D tiled kernel for 3x3 / 5x5, 3-D z-marching
            # kernel for 3x3x3 -- the latter also takes split (hi, lo) inputs as 2*C channels with the weights repeated
            tiled3d = dims == 3 and d['last_w'].shape[1] == 32 and d['last_k'] == (3, 3, 3)
            if d['last_strict']:
                last_simt = LAST_MODE != 'tc' and tiled3d
            else:
                last_simt = LAST_MODE == 'simt' or (LAST_MODE == 'auto' and d['last_w'].shape[1] == 32 and
                                                    ((dims == 2 and d['last_k'][1] in (3, 5)) or tiled3d))
            if not last_simt:
                y = torch.empty((N, D, H, W), dtype=torch.float32, device=x.device)
                ops.tc_conv(d['last_tc'], [o2], (N, D, H, W), out=None, dot_out=y, dot_affine=denorm_stats, rng=rng)
            else:
                y = ops.conv_last(o2, d['last_c'], d['last_w2'] if d['last_strict'] else d['last_w'], d['last_b'], d['last_k'], 1,
                                  d['last_pad'], stats=denorm_stats, rng=rng)
    if dims == 2:
        return y.view(y.shape[0], 1, y.shape[2], y.shape[3])
    return y.view(y.shape[0], 1, y.shape[1], y.shape[2], y.shape[3])


# ------------------------------------------------------------------------------------------------
# fcnn (DenoiseNet2) and affine denoisers (same kernels, small deltas)
# ------------------------------------------------------------------------------------------------
def _build_fcnn_plan(model, device):
    c0, c1, c2 = model.net[0], model.net[2], model.net[4]
    k = c0.weight.shape[-1]
    nf = c0.weight.shape[0]
    pad = (-(k // 2), -(k // 2), 0)
    ld = _tap_ld(k * k)
    strict = PRECISION == 'strict'
    p0 = ops.pack_tc_conv([ConvPart(c0.weight.detach().reshape(nf, k * k, 1, 1), ld, 1, split=strict)], c0.bias, _rup(nf), 0.1, device, strict=strict)
    p1 = ops.pack_tc_conv([ConvPart(c1.weight, _rup(nf), 1, pad, split=strict)], c1.bias, _rup(c1.weight.shape[0]), 0.1, device, strict=strict)
    one = torch.zeros(1, 1, 1, 1); one[0, 0, 0, 0] = 1.0
    p2 = ops.pack_tc_conv([ConvPart(c2.weight, _rup(c1.weight.shape[0]), 1, pad, split=strict)], None, 16, 1.0, device, dot_w=one,
                          dot_b=float(c2.bias.detach()[0]), strict=strict, split_out=False)
    return dict(k=k, ld=ld, p0=p0, p1=p1, p2=p2, strict=strict)


def fcnn_forward(model, x: torch.Tensor, denorm_stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """DenoiseNet2.forward (reference denoising/models.py:52-66); x: fp32 [N,1,H,W] on the device."""
    ops.require_cuda(x, 'denoiser input')
    plan = _cached(model, 'fcnn', _state_key(model, ('fcnn', str(x.device))), lambda: _build_fcnn_plan(model, x.device))
    xi = x[:, 0].contiguous().float()
    N, H, W = xi.shape
    rng = ops.range_scale(xi) if RANGE_GUARD else None
    col = ops.im2col_first(xi, plan['k'], plan['k'] // 2, plan['ld'], rng=rng, split=plan['strict'])
    h0 = torch.empty((N, 1, H, W, plan['p0'].out_channels), dtype=torch.float16, device=x.device)
    ops.tc_conv(plan['p0'], [col], (N, 1, H, W), out=h0, rng=rng)
    h1 = torch.empty((N, 1, H, W, plan['p1'].out_channels), dtype=torch.float16, device=x.device)
    ops.tc_conv(plan['p1'], [h0], (N, 1, H, W), out=h1, rng=rng)
    y = torch.empty((N, 1, H, W), dtype=torch.float32, device=x.device)
    ops.tc_conv(plan['p2'], [h1], (N, 1, H, W), out=None, dot_out=y, dot_affine=denorm_stats, rng=rng)
    return y


def affine_forward(model, x: torch.Tensor, denorm_stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """AffineDenoise.forward (reference filters.py:40-48): one same-padded 1->1 filter."""
    ops.require_cuda(x, 'denoiser input')
    w = model.filter.weight.detach().to(x.device, torch.float32)[0, 0][None].contiguous()
    y = ops.filter_f32(x[:, 0].contiguous().float()[:, None], w, float(model.filter.bias.detach()[0]))
    if denorm_stats is not None:
        y = ops.affine(y, denorm_stats, inverse=True)
    return y
