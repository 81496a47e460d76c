"""Drop-in mirror of topaz/model/features/basic.py:12-111 (conv31/63/127 stacks: conv -> BN -> PReLU,
stride 2 per layer, fill() turns stride into dilation).  Children are parameter containers; forward runs
the sm_100a kernels via topaz_b200.engine."""
from __future__ import print_function, division
from typing import List

import torch
import torch.nn as nn

from topaz_b200.model.utils import insize_from_outsize


class BasicConv(nn.Module):
    def __init__(self, layers: List[int], units: int, unit_scaling: int = 1, dropout: float = 0,
                 bn: bool = True, pooling=None, activation=nn.PReLU, dims: int = 2):
        super().__init__()
        if dims not in (2, 3):
            raise ValueError(f'Unsupported number of dimensions: {dims}. Try dims=2 or dims=3.')
        if pooling is not None:
            raise NotImplementedError('topaz_b200: pooled conv31/63/127 extractors are outside the B200 hot path')
        conv = nn.Conv2d if dims == 2 else nn.Conv3d
        batch_norm = nn.BatchNorm2d if dims == 2 else nn.BatchNorm3d
        use_bias = (not bn)
        stride = 2
        sizes = layers
        layers, strides = [], []
        nin = 1
        for size in sizes[:-1]:
            layers += [conv(nin, units, size, stride=stride, bias=use_bias)]
            strides += [stride]
            if bn:
                layers += [batch_norm(units)]
                strides += [1]
            layers += [activation()]
            strides += [1]
            if dropout > 0:
                layers += [nn.Dropout(p=dropout)]
            nin = units
            units *= unit_scaling
        layers += [conv(nin, units, sizes[-1], bias=use_bias)]
        strides += [1]
        if bn:
            layers += [batch_norm(units)]
            strides += [1]
        layers += [activation()]
        if dropout > 0:
            layers += [nn.Dropout(p=dropout)]
        strides += [1]
        self.strides = strides
        self.width = insize_from_outsize(layers, 1)
        self.filled = False
        self.features = nn.Sequential(*layers)
        self.latent_dim = units
        self.dims = dims

    def fill(self, stride: int = 1):
        for mod, mod_stride in zip(self.features.children(), self.strides):
            if hasattr(mod, 'dilation'):
                mod.dilation = tuple(stride for _ in range(self.dims))
            if hasattr(mod, 'stride'):
                mod.stride = tuple(1 for _ in range(self.dims))
            stride *= mod_stride
        self.filled = True
        return stride

    def unfill(self):
        for mod, mod_stride in zip(self.features.children(), self.strides):
            if hasattr(mod, 'dilation'):
                mod.dilation = tuple(1 for _ in range(self.dims))
            if hasattr(mod, 'stride'):
                mod.stride = tuple(mod_stride for _ in range(self.dims))
        self.filled = False

    def forward(self, x):
        from topaz_b200 import engine
        return engine.features_forward(self, x)


class Conv127(BasicConv):
    def __init__(self, units: int, **kwargs):
        super().__init__([7, 5, 5, 5, 5], units, dims=2, **kwargs)


class Conv63(BasicConv):
    def __init__(self, units: int, **kwargs):
        super().__init__([7, 5, 5, 5], units, dims=2, **kwargs)
