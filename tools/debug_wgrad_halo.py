"""Per-tap error of the halo-resident wgrad kernel (MN-major tcgen05 operands) against torch fp32 on small cases."""
import ctypes as C
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from topaz_b200 import _lib

L = _lib.lib()
P = lambda t: C.c_void_p(t.data_ptr())
for (N, H, dil, org) in [(1, 9, 1, 0), (4, 33, 1, 0), (40, 31, 2, 0), (3, 20, 1, 2)]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, 32, H, H, generator=g)
    xin = x[:, :, org:, org:]
    Ho = xin.shape[2] - 2 * dil
    xin = xin[:, :, :Ho + 2 * dil, :Ho + 2 * dil]
    dy = torch.randn(N, 32, Ho, Ho, generator=g)
    gw = torch.nn.grad.conv2d_weight(xin.contiguous(), (32, 32, 3, 3), dy, dilation=dil)
    xd = x.permute(0, 2, 3, 1).contiguous().cuda(); dyd = dy.permute(0, 2, 3, 1).contiguous().cuda()
    dw = torch.zeros(32, 32, 3, 3, device='cuda')
    rc = L.tpz_conv_wgrad_tc(P(xd), N, H, H, 32, P(dyd), Ho, Ho, 32, 3, 3, 1, dil, org, P(dw), None)
    torch.cuda.synchronize()
    d = (dw.cpu() - gw).abs()
    scale = gw.abs().max()
    print(f'N={N} H={H} dil={dil} org={org} rc={rc}: max rel err {float(d.max() / scale):.3e}; per tap:')
    print(np.array2string((d.amax(dim=(0, 1)) / scale).numpy(), precision=2))
