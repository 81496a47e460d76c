"""Drop-in mirrors of topaz/model/features/resnet.py (BasicConv :50-105, ResidA :108-204, ResNet :208-251,
ResNet8 :280-306, ResNet16 :309-339): same constructor arguments, attributes (kernel_size, stride, dilation,
og_dilation, padding, width, latent_dim, pad), state_dict keys and fill()/unfill() semantics.

The torch.nn.Conv*/BatchNorm* children are parameter containers only (so old ``.sav`` state_dicts load);
their forward() is never used: ``forward`` runs the sm_100a kernels through ``topaz_b200.engine``.
"""
from __future__ import division, print_function

import torch
import torch.nn as nn

from topaz_b200.model.utils import EngineStateMixin, insize_from_outsize


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


def _nd(value, dims):
    return (value,) * dims


_CONVS = {2: nn.Conv2d, 3: nn.Conv3d}
_NORMS = {2: nn.BatchNorm2d, 3: nn.BatchNorm3d}


def _check_dims(dims):
    if dims not in _CONVS:
        raise ValueError(f'Unsupported number of dimensions: {dims}. Try dims=2 or dims=3.')


class BasicConv(nn.Module):
    """conv (+BN) + activation block of the ResNets (reference resnet.py:50-105).  Attributes mirror the reference:
    kernel_size, stride (training stride), dilation (current), og_dilation, padding, dims; children conv / bn / act."""

    def __init__(self, nin, nout, kernel_size, dilation=1, stride=1, bn=False, activation=nn.ReLU, dims=2):
        super().__init__()
        _check_dims(dims)
        self.conv = _CONVS[dims](nin, nout, kernel_size, dilation=dilation, stride=stride, bias=not bn)
        if bn:
            self.bn = _NORMS[dims](nout)
        self.act = activation(inplace=True)
        self.kernel_size, self.stride, self.dims = kernel_size, stride, dims
        self.dilation = self.og_dilation = dilation
        self.padding = 0

    def set_padding(self, pad):
        self.padding = self.dilation * (self.kernel_size // 2) if pad else 0
        self.conv.padding = _nd(self.padding, self.dims)

    def fill(self, stride):
        """Dense mode: this conv runs at stride 1 with its dilation multiplied by the cumulative stride so far."""
        c = self.conv
        c.dilation, c.stride = _nd(self.og_dilation * stride, self.dims), _nd(1, self.dims)
        c.padding = tuple(p * stride for p in c.padding)
        self.dilation = self.dilation * stride
        return self.stride

    def unfill(self):
        c = self.conv
        factor = self.dilation // self.og_dilation
        c.dilation, c.stride = _nd(self.og_dilation, self.dims), _nd(self.stride, self.dims)
        c.padding = tuple(p // factor for p in c.padding)
        self.dilation = self.og_dilation


class ResidA(nn.Module):
    """Residual block (reference resnet.py:108-204): conv0 3x3 -> act -> conv1 3x3 (dilated, optionally strided);
    the block input, centre-cropped by conv0.dilation + conv1.dilation, is added (through a bias-free 1x1 ``proj``
    when the channel count changes) before the optional BN and the final activation."""

    def __init__(self, nin, nhidden, nout, dilation=1, stride=1, activation=nn.ReLU, bn=False, dims=2):
        super().__init__()
        _check_dims(dims)
        conv, norm = _CONVS[dims], _NORMS[dims]
        self.dims, self.bn = dims, bn
        if nin != nout:
            self.proj = conv(nin, nout, 1, stride=stride, bias=False)
        self.conv0 = conv(nin, nhidden, 3, bias=not bn)
        if bn:
            self.bn0 = norm(nhidden)
        self.act0 = activation(inplace=True)
        self.conv1 = conv(nhidden, nout, 3, dilation=dilation, stride=stride, bias=not bn)
        if bn:
            self.bn1 = norm(nout)
        self.act1 = activation(inplace=True)
        self.kernel_size = 3 + 2 * dilation       # receptive field of the block at stride 1
        self.stride, self.dilation, self.padding = stride, 1, 0

    def _skip_convs(self):
        return [self.proj] if hasattr(self, 'proj') else []

    def fill(self, stride):
        self.conv0.dilation = _nd(stride, self.dims)
        self.conv1.dilation = tuple(d * stride for d in self.conv1.dilation)
        for c in [self.conv1] + self._skip_convs():
            c.stride = _nd(1, self.dims)
        self.dilation *= stride
        return self.stride

    def unfill(self):
        self.conv0.dilation = _nd(1, self.dims)
        self.conv1.dilation = tuple(d // self.dilation for d in self.conv1.dilation)
        for c in [self.conv1] + self._skip_convs():
            c.stride = _nd(self.stride, self.dims)
        self.dilation = 1


class ResNet(EngineStateMixin, nn.Module):
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
