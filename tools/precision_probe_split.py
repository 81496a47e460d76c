#!/usr/bin/env python
"""Follow-up of tools/precision_probe.py for unet-3d-10a: how much of each late layer's 11-bit operand error comes from the
WEIGHT rounding and how much from the ACTIVATION rounding, and what cheaper variants of the `auto` precision mode would cost in
parity (a layer that splits only its activations needs 2 k-blocks per (tap, chunk) instead of 3).  CPU only (oracle + goldens).
Result (32^3 block of the packaged weights; max-rel / rel-L2 against fp32):
    current auto (last four convs exact)             4.5e-4 / 4.0e-4
    dec1.0 with rounded weights, split activations   6.2e-4 / 6.7e-4     (dec1.0 is half of the network's FLOPs)
    dec1.0 with split weights, rounded activations   9.9e-4 / 7.7e-4
    dec1.0 fast, the other three exact               1.1e-3 / 9.5e-4
    dec2.2 fast, the other three exact               1.3e-3 / 1.4e-3
so every one of the four layers is needed, and the only saving inside the 1e-3 bound (dec1.0 without its x_hi*w_lo blocks, about
-14 % of the patch time) leaves too little margin at 192^3 (auto measures 5.5e-4 L2 there) to adopt without a hardware check."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from common import gold, weights_of, rel_err
from oracle import topaz_oracle as O
g = gold('unet3d_pretrained_10a'); sd = weights_of(g)
x = g['x32']
orig = O._conv
ref = O.unet_forward(sd, x).numpy()
names = [k[:-7] for k in sd if k.endswith('.weight')]
def is_layer(w, n):
    return w.shape == sd[n + '.weight'].shape and torch.equal(w, torch.from_numpy(sd[n + '.weight']))
def run(cfg):
    """cfg: name -> (round_w, round_a)"""
    def conv(x_, w, b=None, stride=1, dilation=1, padding=0):
        for n, (rw, ra) in cfg.items():
            if is_layer(w, n):
                if rw: w = w.half().float()
                if ra: x_ = x_.half().float()
        return orig(x_, w, b, stride, dilation, padding)
    O._conv = conv
    try:
        return rel_err(O.unet_forward(sd, x).numpy(), ref)
    finally:
        O._conv = orig
fmt = lambda t: f'{t[0]:.2e} / {t[1]:.2e}'
for n in ['dec2.0', 'dec2.2', 'dec1.0', 'dec1.2', 'dec1.4']:
    print(n, 'w only', fmt(run({n: (True, False)})), ' a only', fmt(run({n: (False, True)})))
allr = {n: (True, True) for n in names}
def variant(**over):
    c = dict(allr); c.update({k.replace('_', '.'): v for k, v in over.items()}); return c
print('current auto  :', fmt(run(variant(dec2_2=(False, False), dec1_0=(False, False), dec1_2=(False, False), dec1_4=(False, False)))))
for desc, ov in [
    ('a-split only on the four', dict(dec2_2=(True, False), dec1_0=(True, False), dec1_2=(True, False), dec1_4=(True, False))),
    ('w-split only on the four', dict(dec2_2=(False, True), dec1_0=(False, True), dec1_2=(False, True), dec1_4=(False, True))),
    ('dec1.0 a-only, rest full', dict(dec2_2=(False, False), dec1_0=(True, False), dec1_2=(False, False), dec1_4=(False, False))),
    ('dec1.0 w-only, rest full', dict(dec2_2=(False, False), dec1_0=(False, True), dec1_2=(False, False), dec1_4=(False, False))),
    ('dec1.0 fast, rest full', dict(dec2_2=(False, False), dec1_2=(False, False), dec1_4=(False, False))),
    ('dec2.2 fast, rest full', dict(dec1_0=(False, False), dec1_2=(False, False), dec1_4=(False, False))),
]:
    print(f'{desc:28s}:', fmt(run(variant(**ov))))
