"""Drop-in mirrors of topaz/model/features/resnet.py (BasicConv :50-105, ResidA :108-204, ResNet :208-251,
ResNet8 :280-306, ResNet16 :309-339): same constructor arguments, attributes (kernel_size, stride, dilation,
og_dilation, padding, width, latent_dim, pad), state_dict keys and fill()/unfill() semantics.

The torch.nn.Conv*/BatchNorm* children are parameter containers only (so old ``.sav`` state_dicts load);
their forward() is never used: ``forward`` runs the sm_100a kernels through ``topaz_b200.engine``.
"""
from __future__ import division, print_function

import torch
import torch.nn as nn

from topaz_b200.model.utils import insize_from_outsize


class MaxPool(nn.Module):
    """resnet.py:10-47.  Only reachable with --pooling max (non-default); kept for API parity."""
    def __init__(self, kernel_size, stride=1, dims=2):
        super().__init__()
        self.kernel_size, self.stride, self.og_stride = kernel_size, stride, stride
        self.dilation, self.padding, self.dims = 1, 0, dims

    def fill(self, stride):
        self.dilation, self.stride = stride, 1
        return self.og_stride

    def unfill(self):
        self.dilation, self.stride = 1, self.og_stride

    def forward(self, x):
        raise NotImplementedError('topaz_b200: pooling="max" feature extractors are outside the B200 hot path')


class BasicConv(nn.Module):
    def __init__(self, nin, nout, kernel_size, dilation=1, stride=1, bn=False, activation=nn.ReLU, dims=2):
        super().__init__()
        if dims not in (2, 3):
            raise ValueError(f'Unsupported number of dimensions: {dims}. Try dims=2 or dims=3.')
        conv = nn.Conv2d if dims == 2 else nn.Conv3d
        batch_norm = nn.BatchNorm2d if dims == 2 else nn.BatchNorm3d
        self.conv = conv(nin, nout, kernel_size, dilation=dilation, stride=stride, bias=(not bn))
        if bn:
            self.bn = batch_norm(nout)
        self.act = activation(inplace=True)
        self.kernel_size = kernel_size
        self.stride = stride
        self.dilation = dilation
        self.og_dilation = dilation
        self.padding = 0
        self.dims = dims

    def set_padding(self, pad):
        p = self.dilation * (self.kernel_size // 2) if pad else 0
        self.conv.padding = tuple(p for _ in range(self.dims))
        self.padding = p

    def fill(self, stride):
        self.conv.dilation = tuple(self.og_dilation * stride for _ in range(self.dims))
        self.conv.stride = tuple(1 for _ in range(self.dims))
        self.conv.padding = tuple(pad * stride for pad in self.conv.padding)
        self.dilation *= stride
        return self.stride

    def unfill(self):
        stride = self.dilation // self.og_dilation
        self.conv.dilation = tuple(self.og_dilation for _ in range(self.dims))
        self.conv.stride = tuple(self.stride for _ in range(self.dims))
        self.conv.padding = tuple(pad // stride for pad in self.conv.padding)
        self.dilation = self.og_dilation


class ResidA(nn.Module):
    def __init__(self, nin, nhidden, nout, dilation=1, stride=1, activation=nn.ReLU, bn=False, dims=2):
        super().__init__()
        if dims not in (2, 3):
            raise ValueError(f'Unsupported number of dimensions: {dims}. Try dims=2 or dims=3.')
        self.dims = dims
        conv = nn.Conv2d if dims == 2 else nn.Conv3d
        batch_norm = nn.BatchNorm2d if dims == 2 else nn.BatchNorm3d
        self.bn = bn
        bias = (not bn)
        if nin != nout:
            self.proj = conv(nin, nout, 1, stride=stride, bias=False)
        self.conv0 = conv(nin, nhidden, 3, bias=bias)
        if self.bn:
            self.bn0 = batch_norm(nhidden)
        self.act0 = activation(inplace=True)
        self.conv1 = conv(nhidden, nout, 3, dilation=dilation, stride=stride, bias=bias)
        if self.bn:
            self.bn1 = batch_norm(nout)
        self.act1 = activation(inplace=True)
        self.kernel_size = 2 * dilation + 3
        self.stride = stride
        self.dilation = 1
        self.padding = 0

    def fill(self, stride):
        self.conv0.dilation = tuple(stride for _ in range(self.dims))
        self.conv1.dilation = tuple(dil * stride for dil in self.conv1.dilation)
        self.conv1.stride = tuple(1 for _ in range(self.dims))
        if hasattr(self, 'proj'):
            self.proj.stride = tuple(1 for _ in range(self.dims))
        self.dilation = self.dilation * stride
        return self.stride

    def unfill(self):
        self.conv0.dilation = tuple(1 for _ in range(self.dims))
        self.conv1.dilation = tuple(dil // self.dilation for dil in self.conv1.dilation)
        self.conv1.stride = tuple(self.stride for _ in range(self.dims))
        if hasattr(self, 'proj'):
            self.proj.stride = tuple(self.stride for _ in range(self.dims))
        self.dilation = 1


class ResNet(nn.Module):
    '''ResNet utility functions. Must be subclassed to define network architecture.'''
    def __init__(self, dims=2, **kwargs):
        super().__init__()
        self.dims = dims
        if 'pooling' in kwargs and kwargs['pooling'] == 'max':
            kwargs['pooling'] = MaxPool
        modules = self.make_modules(**kwargs)
        self.features = nn.Sequential(*modules)
        self.width = insize_from_outsize(modules, 1)
        self.pad = False

    def fill(self, stride=1):
        for mod in self.features.children():
            if hasattr(mod, 'fill'):
                stride *= mod.fill(stride)
        self.pad = True
        return stride

    def unfill(self):
        for mod in self.features.children():
            if hasattr(mod, 'unfill'):
                mod.unfill()
        self.pad = False

    def set_padding(self, pad):
        self.pad = pad

    def forward(self, x):
        from topaz_b200 import engine
        return engine.features_forward(self, x)


def _units(units):
    if units is None:
        return [32, 64, 128]
    if type(units) is not list:
        units = int(units)
        return [units, 2 * units, 4 * units]
    return units


# Architecture tables.  Entries: ('conv', out_level, kernel, strided) | ('resid', in_level, out_level,
# dilation, strided) | ('pool',) = optional pooling slot (+ dropout slot) exactly where the reference puts them.
_TABLES = {
    'ResNet8': [('conv', 0, 7, True), ('pool',), ('resid', 0, 0, 2, False), ('resid', 0, 1, 2, True), ('pool',),
                ('resid', 1, 1, 2, False), ('conv', 2, 5, False), ('drop',)],
    'ResNet16': [('conv', 0, 7, False), ('resid', 0, 0, 1, True), ('pool',), ('resid', 0, 0, 1, False),
                 ('resid', 0, 0, 1, False), ('resid', 0, 0, 1, False), ('resid', 0, 1, 1, True), ('pool',),
                 ('resid', 1, 1, 1, False), ('resid', 1, 1, 1, False), ('conv', 2, 5, False), ('drop',)],
}


class _TableResNet(ResNet):
    def make_modules(self, units=[32, 64, 128], bn=True, dropout=0.0, activation=nn.ReLU, pooling=None, **kwargs):
        units = _units(units)
        self.num_features = self.latent_dim = units[-1]
        self.stride = 1 if pooling is not None else 2      # strided convs replace pooling when pooling is None
        mods, nin = [], 1
        for entry in _TABLES[type(self).__name__]:
            kind = entry[0]
            if kind == 'conv':
                _, lvl, k, strided = entry
                mods.append(BasicConv(nin, units[lvl], k, stride=self.stride if strided else 1, bn=bn,
                                      activation=activation, dims=self.dims))
                nin = units[lvl]
            elif kind == 'resid':
                _, li, lo, dil, strided = entry
                mods.append(ResidA(units[li], units[li], units[lo], dilation=dil,
                                   stride=self.stride if strided else 1, bn=bn, activation=activation,
                                   dims=self.dims))
                nin = units[lo]
            else:
                if kind == 'pool' and pooling is not None:
                    mods.append(pooling(3, stride=2, dims=self.dims))
                if dropout > 0:
                    mods.append(nn.Dropout(p=dropout))
        return mods


class ResNet8(_TableResNet):
    """resnet.py:280-306: conv7(s2) | ResidA(d2) | ResidA(d2, s2, widen) | ResidA(d2) | conv5; width 71."""


class ResNet16(_TableResNet):
    """resnet.py:309-339: conv7 | ResidA(s2) | 3x ResidA | ResidA(s2, widen) | 2x ResidA | conv5."""
