#!/usr/bin/env python
"""Golden vectors for the BatchNorm training path (`topaz train` default: --bn on), produced by the REAL reference
imported read-only from /root/reference (h5py stub).  Build container only:
    PYTHONDONTWRITEBYTECODE=1 python tools/make_goldens_bn.py
Writes tests/golden/ge_binomial_u32_bn.npz: 3 GE_binomial steps of ResNet8(units=32, bn=True) from seeded weights
(tests/common.seeded_state, seed 401), gradients of step 1, every parameter / buffer after step 3, the unfilled
eval-mode scores of a crop batch and the filled dense scores of a small image after training."""
import os, sys
sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'tools', 'stubs'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import torch.nn as nn

torch.set_num_threads(os.cpu_count())
from common import seeded_state, GOLD  # noqa: E402
from topaz.model.classifier import LinearClassifier
from topaz.model.features.resnet import ResNet8
from topaz.methods import GE_binomial

rng = np.random.default_rng
SEED = 401
m = LinearClassifier(ResNet8(units=32, bn=True))
sd = seeded_state({k: tuple(v.shape) for k, v in m.state_dict().items()}, SEED)
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
m.train()
optim = torch.optim.Adam(m.parameters(), lr=2e-4)
tr = GE_binomial(m, optim, nn.BCEWithLogitsLoss(), 0.035, l2=0.0, slack=1.0)
B = 64
Y = np.array([1.0] * 4 + [0.0] * (B - 4))
outs, grads1 = [], None
for step in range(3):
    X = rng(4000 + step).standard_normal((B, 71, 71)).astype(np.float32)
    orig = optim.step
    if step == 0:
        def grab(*a, **k):
            global grads1
            grads1 = {n: p.grad.detach().clone().numpy() for n, p in m.named_parameters()}
            return orig(*a, **k)
        optim.step = grab
    outs.append(tr.step(torch.from_numpy(X), torch.from_numpy(Y)))
    optim.step = orig
final = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
m.eval()
Xe = rng(4100).standard_normal((8, 71, 71)).astype(np.float32)
with torch.no_grad():
    y_crops = m(torch.from_numpy(Xe)).numpy()
    m.fill()
    xd = rng(4101).standard_normal((1, 96, 80)).astype(np.float32)
    y_dense = m(torch.from_numpy(xd)).numpy()
    m.unfill()
path = os.path.join(GOLD, 'ge_binomial_u32_bn.npz')
np.savez_compressed(path, seed=np.int64(SEED), B=np.int64(B), Y=Y, pi=np.float64(0.035), outs=np.array(outs, dtype=np.float64),
                    y_crops=y_crops, x_dense=xd, y_dense=y_dense,
                    **{'g1.' + k: v for k, v in grads1.items()}, **{'p3.' + k: v for k, v in final.items()})
print('wrote', path, f'{os.path.getsize(path) / 1e6:.2f} MB', 'outs', outs)

# ---- PReLU extractors in training (`topaz train -m conv31|conv63|conv127`): conv31, 16 units x2, with and without BatchNorm ----
from topaz.model.factory import get_feature_extractor
for tag, bn in (('ge_binomial_conv31_bn', True), ('ge_binomial_conv31_nobn', False)):
    seed = 402 if bn else 403
    m = LinearClassifier(get_feature_extractor('conv31', units=16, dropout=0.0, bn=bn, unit_scaling=2, pooling=None, dims=2))
    sd = seeded_state({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    m.train()
    optim = torch.optim.Adam(m.parameters(), lr=2e-4)
    tr = GE_binomial(m, optim, nn.BCEWithLogitsLoss(), 0.035, l2=0.0, slack=1.0)
    B = 32
    Y = np.array([1.0] * 3 + [0.0] * (B - 3))
    outs, grads1 = [], None
    for step in range(2):
        X = rng(4200 + step).standard_normal((B, m.width, m.width)).astype(np.float32)
        orig = optim.step
        if step == 0:
            def grab(*a, **k):
                global grads1
                grads1 = {n: p.grad.detach().clone().numpy() for n, p in m.named_parameters()}
                return orig(*a, **k)
            optim.step = grab
        outs.append(tr.step(torch.from_numpy(X), torch.from_numpy(Y)))
        optim.step = orig
    final = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    path = os.path.join(GOLD, tag + '.npz')
    np.savez_compressed(path, seed=np.int64(seed), B=np.int64(B), width=np.int64(m.width), Y=Y, pi=np.float64(0.035),
                        outs=np.array(outs, dtype=np.float64), keys=np.array(list(sd.keys())),
                        **{'g1.' + k: v for k, v in grads1.items()}, **{'p2.' + k: v for k, v in final.items()})
    print('wrote', path, f'{os.path.getsize(path) / 1e6:.2f} MB', 'width', m.width, 'outs', outs)

# ---- nn.Dropout in training (`topaz train --dropout`): ResNet8, 16 units, BatchNorm, p = 0.25; one GE_binomial step with the
# keep-masks of the three Dropout layers recorded by forward hooks (torch's mask stream cannot be reproduced outside torch,
# so parity is "same masks -> same logits, loss and gradients") ----
torch.manual_seed(1234)
P_DROP = 0.25
m = LinearClassifier(ResNet8(units=16, bn=True, dropout=P_DROP))
sd = seeded_state({k: tuple(v.shape) for k, v in m.state_dict().items()}, 404)
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
m.train()
masks = []
for mod in m.modules():
    if isinstance(mod, nn.Dropout):
        mod.register_forward_hook(lambda mod, inp, out: masks.append(((out != 0) | (inp[0] == 0)).numpy().copy()))
optim = torch.optim.Adam(m.parameters(), lr=2e-4)
tr = GE_binomial(m, optim, nn.BCEWithLogitsLoss(), 0.035, l2=0.0, slack=1.0)
B = 16
Y = np.array([1.0] * 2 + [0.0] * (B - 2))
X = rng(4300).standard_normal((B, 71, 71)).astype(np.float32)
grads1 = None
orig = optim.step
def grab(*a, **k):
    global grads1
    grads1 = {n: p.grad.detach().clone().numpy() for n, p in m.named_parameters()}
    return orig(*a, **k)
optim.step = grab
out = tr.step(torch.from_numpy(X), torch.from_numpy(Y))
assert len(masks) == 3, len(masks)
path = os.path.join(GOLD, 'ge_binomial_u16_dropout.npz')
np.savez_compressed(path, seed=np.int64(404), B=np.int64(B), Y=Y, pi=np.float64(0.035), p=np.float64(P_DROP),
                    out=np.array(out, dtype=np.float64), keys=np.array(list(sd.keys())),
                    **{f'mask{i}': np.packbits(mk.reshape(-1)) for i, mk in enumerate(masks)},
                    **{f'mask{i}.shape': np.array(mk.shape) for i, mk in enumerate(masks)},
                    **{'g1.' + k: v for k, v in grads1.items()})
print('wrote', path, f'{os.path.getsize(path) / 1e6:.2f} MB', 'out', out, [mk.mean() for mk in masks])
