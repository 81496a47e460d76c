#!/usr/bin/env python
"""Per-layer timing of the tensor-core conv on a large image (device-resident, CUDA events), to separate the
effects of dilation (poly-phase strided TMA), channel counts and kernel variant.  Also prints the MMA-rate probe
(cycles per M=128 MMA vs A-operand alignment).  Bring-up tool, not on the product path."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from topaz_b200 import ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "lab"))
import lab
from topaz_b200.ops import ConvPart


def rate_probe():
    out = {}
    for N in (64, 128, 256):
        for (shift, sbo) in ((0, 8), (1, 8), (4, 8), (0, 10), (3, 10), (0, 16), (5, 12)):
            for two in (False, True):
                c = lab.lab_umma_rate(N, shift, sbo, 2000, two)
                out[f'N{N}_shift{shift}_sbo{sbo}_two{int(two)}'] = c
                print(f'mma rate N={N} shift={shift} sbo_rows={sbo} two_acc={int(two)}: {c:.1f} cycles/MMA (ideal {N/2:.0f})')
    return out


def layer(cin, co, k, dil, S=2048, variant='auto', reps=5, residual_src=False):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 1, S, S, cin, generator=g).half().cuda()
    w = torch.randn(co, cin, k, k, generator=g) * 0.05
    parts = [ConvPart(w, cin, dil)]
    srcs = [x]
    if residual_src:
        e = (k - 1) * dil // 2
        parts.append(ConvPart(torch.eye(co).reshape(co, co, 1, 1), cin, 1, (e, e, 0)))
        srcs.append(x)
    plan = ops.pack_tc_conv(parts, torch.zeros(co), co, 0.0, 'cuda')
    Ho = S - (k - 1) * dil
    out = torch.empty(1, 1, Ho, Ho, co, dtype=torch.float16, device='cuda')
    ops.TC_VARIANT = variant
    for _ in range(2):
        ops.tc_conv(plan, srcs, (1, 1, Ho, Ho), out=out)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        ops.tc_conv(plan, srcs, (1, 1, Ho, Ho), out=out)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / reps
    fl = 2.0 * Ho * Ho * co * cin * k * k
    print(f'layer cin={cin} co={co} k={k} dil={dil} S={S} variant={variant} extra_src={residual_src}: {ms:.3f} ms  {fl/ms/1e9:.0f} TFLOP/s')
    return ms


if __name__ == '__main__':
    rate_probe()
    for variant in ('v2', 'v1'):
        for (cin, co, k, dil) in ((64, 64, 3, 1), (64, 64, 3, 2), (64, 64, 3, 4), (128, 128, 3, 1), (128, 128, 3, 4), (128, 128, 3, 8),
                                  (64, 64, 1, 1), (128, 256, 5, 4)):
            layer(cin, co, k, dil, variant=variant)
