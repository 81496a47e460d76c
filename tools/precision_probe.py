#!/usr/bin/env python
"""Per-layer error budget of 11-bit (fp16 / TF32) operands in the reference's pretrained denoisers, on the CPU.

Runs the oracle with the operands of ONE convolution at a time rounded to fp16 (weights and input activations), then with
all layers rounded except a growing tail, and prints the SURVEY 8(c) metric against the fp32 result.  This is the evidence
behind engine._unet_precision: for unet-3d-10a on N(0,1) input the output (max 0.16, std 0.04) is a ~100x cancellation of
O(1) features, and the last four convolutions carry 90 % of the 5e-3 error; the 2-D unet stays at 5e-4.
    python tools/precision_probe.py [unet3d_pretrained_10a | unet_pretrained]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
from common import gold, weights_of, rel_err
from oracle import topaz_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else 'unet3d_pretrained_10a'
g = gold(name); sd = weights_of(g)
x = g['x32'] if 'x32' in g.files else np.random.default_rng(1).standard_normal((1, 1, 256, 256)).astype(np.float32)
orig = O._conv
ref = O.unet_forward(sd, x).numpy()
print(f'{name}: input {x.shape}, output max {np.abs(ref).max():.3f} std {ref.std():.3f}')
names = [k[:-7] for k in sd if k.endswith('.weight')]


def is_layer(w, n):
    return w.shape == sd[n + '.weight'].shape and torch.equal(w, torch.from_numpy(sd[n + '.weight']))


def run(rounded, rw=True, ra=True):
    def conv(x_, w, b=None, stride=1, dilation=1, padding=0):
        if any(is_layer(w, n) for n in rounded):
            if rw: w = w.half().float()
            if ra: x_ = x_.half().float()
        return orig(x_, w, b, stride, dilation, padding)
    O._conv = conv
    try:
        return rel_err(O.unet_forward(sd, x).numpy(), ref)
    finally:
        O._conv = orig


print('all layers, weights only  (max-rel, rel-L2):', run(names, True, False))
print('all layers, activations only               :', run(names, False, True))
print('all layers, both                           :', run(names))
for n in names:
    print(f'  only {n:8s} {str(sd[n + ".weight"].shape):22s}', run([n]))
tail = []
for n in reversed(names[-4:]):
    tail.append(n)
    print('all layers except', tail, run([m for m in names if m not in tail]))
