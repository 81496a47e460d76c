#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/dp_check.py : data-parallel GE_binomial.step over NCCL must reproduce the
single-process reference golden (3 steps, 64 crops sharded over the ranks).  Prints PASS/FAIL on rank 0."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
from common import gold, weights_of, rel_err

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from topaz_b200.methods import GE_binomial
from topaz_b200.model.factory import get_feature_extractor
from topaz_b200.model.classifier import LinearClassifier
g = gold('ge_binomial_u32'); sd = weights_of(gold('resnet8_u32_pretrained'))
m = LinearClassifier(get_feature_extractor('resnet8', units=32, bn=False))
m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); m.cuda(); m.train()
tr = GE_binomial(m, torch.optim.Adam(m.parameters(), lr=2e-4), nn.BCEWithLogitsLoss(), float(g['pi']))
B = int(g['B']); b = B // world
Y = torch.from_numpy(g['Y'])
outs = []
for step in range(3):
    X = torch.from_numpy(np.random.default_rng(4000 + step).standard_normal((B, 71, 71)).astype(np.float32))
    outs.append(tr.step(X[rank * b:(rank + 1) * b].cuda(), Y[rank * b:(rank + 1) * b].cuda()))
ok = np.allclose(np.array(outs), g['outs'], rtol=1e-3, atol=1e-6)
worst = 0.0
for k, p in m.named_parameters():
    worst = max(worst, max(rel_err(p.detach().cpu().numpy(), g['p3.' + k])))
flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
ref = flat.clone(); dist.broadcast(ref, 0)
same = bool(torch.equal(flat, ref))
if rank == 0:
    print(f'dp_check world={world}: loss tuple match={ok} worst param rel err={worst:.2e} replicas identical={same} ->',
          'PASS' if ok and worst < 1e-3 and same else 'FAIL')
dist.barrier(); dist.destroy_process_group()
