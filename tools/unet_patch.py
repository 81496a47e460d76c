#!/usr/bin/env python
"""One UDenoiseNet forward on a 2048^2 patch (default) or one UDenoiseNet3D forward on a 192^3 patch (`3d`), for
`ncu --metrics gpu__time_duration.sum` launch lists."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from common import gold, weights_of, seeded_state
from common_shapes import unet_shapes
from topaz_b200.denoising.models import UDenoiseNet, UDenoiseNet3D
from topaz_b200.denoise import Denoise, Denoise3D
if len(sys.argv) > 1 and sys.argv[1] == '3d':
    m = UDenoiseNet3D(nf=48, base_width=7, top_width=3)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in weights_of(gold('unet3d_pretrained_10a')).items()})
    dn = Denoise3D(m)
    x = torch.from_numpy(np.random.default_rng(1).standard_normal((1, 192, 192, 192)).astype(np.float32)).cuda()
else:
    m = UDenoiseNet(base_width=11, top_width=5)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in weights_of(gold('unet_pretrained')).items()})
    dn = Denoise(m)
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    x = torch.from_numpy((10 + 3 * np.random.default_rng(1).standard_normal((S, S))).astype(np.float32)).cuda()
for _ in range(3):
    y = dn._denoise_device(x)
torch.cuda.synchronize()
print('ok', tuple(y.shape))
