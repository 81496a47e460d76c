#!/usr/bin/env python
"""Bisect the resnet8_u64 training step: run forward+backward once with the fp32 CUDA-core kernels and once with the
tensor-core (mma.sync 3xTF32) kernels, recording the output of every conv call, and report where they diverge."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from common import gold, weights_of
from topaz_b200 import train_engine as T
from topaz_b200.model.factory import get_feature_extractor
from topaz_b200.model.classifier import LinearClassifier

units = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sd = weights_of(gold(f'resnet8_u{units}_pretrained'))
B = 48
X = torch.from_numpy(np.random.default_rng(77).standard_normal((B, 71, 71)).astype(np.float32)).cuda()
Y = torch.tensor([1.0] * 5 + [0.0] * (B - 5), dtype=torch.float64).cuda()


def run(use_mma):
    T.USE_MMA = use_mma
    m = LinearClassifier(get_feature_extractor('resnet8', units=units, bn=False))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m.cuda(); m.train()
    log = []
    orig = {n: getattr(T, n) for n in ('_conv_fwd', '_conv_dgrad', '_conv_wgrad', '_crop_add', '_relu_bwd')}

    def wrap(name):
        f = orig[name]
        def g(*a, **k):
            r = f(*a, **k)
            if name == '_conv_wgrad':
                out = a[2]
            elif name in ('_crop_add', '_relu_bwd'):
                out = a[0]
            else:
                out = r
            desc = f'{name} ' + ' '.join(str(tuple(t.shape)) if torch.is_tensor(t) else str(t) for t in a[:2]) + f' args={[v for v in a[2:] if not torch.is_tensor(v)]} {k.keys() and list(k.keys())}'
            log.append((desc, out.detach().clone()))
            return r
        return g
    for n in orig:
        setattr(T, n, wrap(n))
    try:
        T.flat_params(m)
        score = m(X).view(-1)
        ds = torch.empty(B, device='cuda'); o5 = torch.empty(5, device='cuda')
        T.ge_loss_grad(score.contiguous(), Y, 0.05, 1.0, 0, B, ds, o5)
        T.backward(m, ds)
    finally:
        for n, f in orig.items():
            setattr(T, n, f)
    return log


a, b = run(False), run(True)
print(len(a), len(b))
for i, ((da, ta), (db, tb)) in enumerate(zip(a, b)):
    err = float((ta - tb).abs().max() / (ta.abs().max() + 1e-30))
    flag = '  <<<<' if err > 1e-4 else ''
    print(f'{i:3d} {err:.2e} {da[:150]}{flag}')
