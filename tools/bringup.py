#!/usr/bin/env python
"""GPU bring-up script (run on the B200 box): staged checks of the tcgen05 path with diagnostics.
Writes gpurun_out/bringup.log and gpurun_out/lab.json.  Never part of the product path."""
import json, os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import torch.nn.functional as F
from topaz_b200 import ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "lab"))
import lab
from topaz_b200.ops import ConvPart

os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
LOG = open(os.path.join(ROOT, 'gpurun_out', 'bringup.log'), 'w')


def log(*a):
    s = ' '.join(str(x) for x in a)
    print(s); LOG.write(s + '\n'); LOG.flush()


def lab():
    res = {}
    g = torch.Generator().manual_seed(0)
    for kc in (64, 32):
        A = torch.randn(512, kc, generator=g).half()
        B = torch.randn(64, kc, generator=g).half()
        Ad, Bd = A.cuda(), B.cuda()
        for sbo in (8, 10, 12):
            for shift in (0, 1, 3, 5, 13):
                rows = torch.tensor([shift + (m // 8) * sbo + (m % 8) for m in range(128)])
                exp = A[rows].float() @ B.float().t()
                D = lab.lab_umma(Ad, Bd, shift, sbo, 0, kc); torch.cuda.synchronize()
                err = (D.cpu() - exp).abs().max().item() / exp.abs().max().item()
                res[f'kc{kc}_sbo{sbo}_shift{shift}'] = err
                log(f'lab kc={kc} sbo={sbo} shift={shift}: rel err {err:.2e}', 'OK' if err < 1e-3 else 'FAIL')
    # TMA element-stride probe: which rows land in smem?
    A = torch.arange(600).float()[:, None].repeat(1, 64).half()      # row r holds the value r
    for stride, start, nrows in ((1, 5, 20), (2, 3, 20), (4, 7, 40), (8, 2, 30)):
        raw = lab.lab_tma_stride(A.cuda(), start, stride, nrows); torch.cuda.synchronize()
        got = raw.cpu().float()[:, 0].tolist()
        exp = [float(start + i * stride) for i in range(nrows)]
        ok = got == exp
        res[f'tma_stride{stride}'] = ok
        log(f'lab tma stride={stride} start={start}: rows {got[:8]}... expected {exp[:8]}...', 'OK' if ok else 'FAIL')
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'lab.json'), 'w'), indent=1)
    return res


def ref_conv(srcs, parts_w, dils, orgs, bias, slope, out_shape, res=None, res_org=(0, 0, 0)):
    """fp32 CPU reference on fp16-rounded operands.  srcs NDHWC fp16 (cpu)."""
    N, Do, Ho, Wo = out_shape
    acc = None
    for x, w, dil, org in zip(srcs, parts_w, dils, orgs):
        w = w.half().float()
        if w.dim() == 4:
            w = w[:, :, None]
        co, ci, kd, kh, kw = w.shape
        xi = x[..., :ci].float().permute(0, 4, 1, 2, 3)
        # window start = org; pad so that index 0 of padded = org (org <= 0) or crop (org > 0)
        P = 64
        xp = F.pad(xi, (P, P, P, P, P, P))
        ox, oy, oz = org
        need = lambda o, n, k: slice(P + o, P + o + n + (k - 1) * dil)
        xp = xp[:, :, need(oz, Do, kd), need(oy, Ho, kh), need(ox, Wo, kw)]
        y = F.conv3d(xp, w, None, dilation=dil)
        acc = y if acc is None else acc + y
    y = acc + bias.view(1, -1, 1, 1, 1)
    if res is not None:
        r = res[:, res_org[2]:res_org[2] + Do, res_org[1]:res_org[1] + Ho, res_org[0]:res_org[0] + Wo].float()
        y = y + r.permute(0, 4, 1, 2, 3)[:, :y.shape[1]]
    y = torch.where(y > 0, y, y * slope)
    return y.permute(0, 2, 3, 4, 1)     # NDHWC fp32


def tc_case(name, N, D, H, W, cins, co, k, dil, pad, slope=0.0, kd=1, residual=False, dot=False, seed=0, KC=None):
    g = torch.Generator().manual_seed(seed)
    c_stores = [(c + 31) // 32 * 32 for c in cins]
    srcs = []
    for c, cs in zip(cins, c_stores):
        x = torch.zeros(N, D, H, W, cs)
        x[..., :c] = torch.randn(N, D, H, W, c, generator=g)
        srcs.append(x.half())
    ws = [torch.randn(co, c, kd, k, k, generator=g) * (2.0 / (sum(cins) * k * k * kd)) ** 0.5 for c in cins]
    bias = 0.1 * torch.randn(co, generator=g)
    org = (-pad, -pad, -pad if kd > 1 else 0)
    Do = D + 2 * pad - (kd - 1) * dil if kd > 1 else D
    Ho, Wo = H + 2 * pad - (k - 1) * dil, W + 2 * pad - (k - 1) * dil
    co_store = (co + 31) // 32 * 32
    dw = torch.randn(co, generator=g) if dot else None
    plan = ops.pack_tc_conv([ConvPart(w, cs, dil, org) for w, cs in zip(ws, c_stores)], bias, co_store, slope, 'cuda',
                            KC=KC, dot_w=dw, dot_b=0.25)
    res = None
    if residual:
        res = torch.zeros(N, Do, Ho, Wo, co_store); res[..., :co] = torch.randn(N, Do, Ho, Wo, co, generator=g); res = res.half()
    out = torch.zeros(N, Do, Ho, Wo, co_store, dtype=torch.float16, device='cuda')
    dout = torch.zeros(N, Do, Ho, Wo, dtype=torch.float32, device='cuda') if dot else None
    ops.tc_conv(plan, [s.cuda() for s in srcs], (N, Do, Ho, Wo), out=out, res=res.cuda() if residual else None,
                dot_out=dout)
    torch.cuda.synchronize()
    exp = ref_conv(srcs, ws, [dil] * len(cins), [org] * len(cins), bias, slope, (N, Do, Ho, Wo), res)
    got = out.cpu().float()[..., :co]
    err = (got - exp).abs().max().item() / max(exp.abs().max().item(), 1e-9)
    msg = f'tc {name}: out {tuple(got.shape)} nkb={len(plan.kblocks)} KC={plan.KC} rel err {err:.3e}'
    if dot:
        expd = (exp * dw).sum(-1) + 0.25
        errd = (dout.cpu() - expd).abs().max().item() / expd.abs().max().item()
        msg += f' dot err {errd:.3e}'
        err = max(err, errd)
    pad_ok = bool((out.cpu()[..., co:] == 0).all()) if co_store > co else True
    log(msg + ('' if pad_ok else ' PAD-CHANNELS-NONZERO') + ('  OK' if err < 2e-3 and pad_ok else '  FAIL'))
    if err >= 2e-3:
        d = (got - exp).abs()
        bad = (d > 2e-3 * exp.abs().max()).nonzero()
        log('   first bad idx:', bad[:8].tolist(), ' n_bad', len(bad), 'of', d.numel())
        log('   got[0,0,0,:4,:4]', got[0, 0, 0, :4, :4].tolist())
        log('   exp[0,0,0,:4,:4]', exp[0, 0, 0, :4, :4].tolist())
    return err


def main():
    log('device', torch.cuda.get_device_name(0))
    stages = [
        ('lab', lab),
        ('1x1 64->64', lambda: tc_case('1x1_64_64', 1, 1, 16, 32, [64], 64, 1, 1, 0)),
        ('3x3 64->64 d1', lambda: tc_case('3x3_64_64', 1, 1, 40, 48, [64], 64, 3, 1, 0)),
        ('3x3 64->64 d2 odd size', lambda: tc_case('3x3d2_odd', 2, 1, 37, 53, [64], 64, 3, 2, 0)),
        ('3x3 same-pad lrelu', lambda: tc_case('3x3_same', 1, 1, 33, 47, [64], 64, 3, 1, 1, slope=0.1)),
        ('KC32 3x3 32->32', lambda: tc_case('kc32', 1, 1, 30, 40, [32], 32, 3, 1, 1)),
        ('96->96 KC32', lambda: tc_case('c96', 1, 1, 30, 40, [96], 96, 3, 1, 1, slope=0.1)),
        ('two sources 96+48 -> 96', lambda: tc_case('2src', 1, 1, 30, 40, [96, 48], 96, 3, 1, 1, slope=0.1)),
        ('5x5 d4 128->256 + dot', lambda: tc_case('conv5dot', 1, 1, 48, 56, [128], 256, 5, 4, 0, dot=True)),
        ('3x3 d8 128->128 + residual', lambda: tc_case('resid', 1, 1, 50, 60, [128], 128, 3, 8, 0, residual=True)),
        ('3d 3x3x3 64->64', lambda: tc_case('3d', 1, 12, 20, 24, [64], 64, 3, 1, 1, kd=3, slope=0.1)),
        ('16-ch out', lambda: tc_case('co16', 1, 1, 20, 24, [32], 16, 3, 1, 1)),
        ('big', lambda: tc_case('big', 1, 1, 300, 420, [64], 128, 3, 4, 0)),
    ]
    runs = [('lab', stages[0][1], 'auto')]
    for name, fn in stages[1:]:
        runs.append((name + ' [v2]', fn, 'v2'))
    runs.append(('big [v1]', stages[-1][1], 'v1'))
    for name, fn, variant in runs:
        t0 = time.time()
        ops.TC_VARIANT = variant
        try:
            fn()
        except Exception:
            log(f'STAGE {name} raised:\n' + traceback.format_exc())
            try:
                torch.cuda.synchronize()
            except Exception as e:
                log('CUDA context is dead:', e)
                break
        log(f'-- stage {name} done in {time.time()-t0:.1f}s')


if __name__ == '__main__':
    main()
