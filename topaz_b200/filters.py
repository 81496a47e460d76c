"""Drop-in mirror of topaz.filters.GaussianDenoise (reference filters.py:6-19, 51-80): normalised Gaussian, truncated at
scale*sigma, applied as a same-padded 1->1 convolution (2-D or 3-D) on the GPU (tpz_filter_f32)."""
import numpy as np
import torch
from torch import nn

from topaz_b200 import ops


def gaussian_filter(sigma, s=11, dims=2):
    dim = s // 2
    ranges = np.arange(-dim, dim + 1)
    grids = np.meshgrid(*([ranges] * dims))
    d = sum(g ** 2 for g in grids)
    return np.exp(-0.5 * d / sigma ** 2)


class GaussianDenoise(nn.Module):
    ''' Apply Gaussian filter with sigma to image. Truncates the kernel at scale times sigma pixels. '''
    def __init__(self, sigma, scale=5, dims=2, use_cuda=True):
        super().__init__()
        width = 1 + 2 * int(np.ceil(sigma * scale))
        f = gaussian_filter(sigma, s=width, dims=dims)
        f /= f.sum()
        self.filter = nn.Conv2d(1, 1, width, padding=width // 2) if dims == 2 else nn.Conv3d(1, 1, width, padding=width // 2)
        self.filter.weight.data[:] = torch.from_numpy(f).float()
        self.filter.bias.data.zero_()
        self.dims = dims
        self.use_cuda = use_cuda

    def forward(self, x):
        """x: [N,1,(D),H,W] fp32 on the device."""
        ops.require_cuda(x, 'filter input')
        w = self.filter.weight.detach().to(x.device, torch.float32)[0, 0]
        xi = x[:, 0].contiguous().float()
        if self.dims == 2:
            xi, w = xi[:, None], w[None]
        y = ops.filter_f32(xi, w.contiguous(), float(self.filter.bias.detach()[0]))
        return y if self.dims == 2 else y[:, None]      # 2-D: [N,1(D),H,W] already is [N,1,H,W]

    @torch.no_grad()
    def apply(self, x):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda().unsqueeze(0).unsqueeze(0)
        return self.forward(x).squeeze().cpu().numpy()
