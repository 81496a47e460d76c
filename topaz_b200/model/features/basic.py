"""conv31 / conv63 / conv127 feature extractors: drop-in for ``topaz.model.features.basic.BasicConv`` (reference
basic.py:12-111).  A stack of valid convolutions (first kernels stride 2, last stride 1), each followed by optional
BatchNorm and an activation (PReLU with one slope by default).  ``fill()`` turns the strides into dilations so the
patch classifier can be evaluated densely; ``unfill()`` restores the training geometry.

The torch.nn children only carry parameters / buffers under the reference's state_dict keys (``features.<i>.*``);
arithmetic runs in the sm_100a kernels through ``topaz_b200.engine``.
"""
from typing import List, Sequence

import torch.nn as nn

from topaz_b200.model.utils import EngineStateMixin, insize_from_outsize

_CONV = {2: nn.Conv2d, 3: nn.Conv3d}
_NORM = {2: nn.BatchNorm2d, 3: nn.BatchNorm3d}


class BasicConv(EngineStateMixin, nn.Module):
    def __init__(self, layers: List[int], units: int, unit_scaling: int = 1, dropout: float = 0, bn: bool = True,
                 pooling=None, activation=nn.PReLU, dims: int = 2):
        super().__init__()
        if dims not in _CONV:
            raise ValueError(f'Unsupported number of dimensions: {dims}. Try dims=2 or dims=3.')
        if pooling is not None:
            raise NotImplementedError('topaz_b200: pooled conv31/63/127 extractors are outside the B200 hot path')
        kernel_sizes: Sequence[int] = list(layers)
        mods, per_module_stride = [], []
        width_in, width_out = 1, units
        for pos, ksize in enumerate(kernel_sizes):
            final = pos == len(kernel_sizes) - 1
            step = 1 if final else 2
            block = [(_CONV[dims](width_in, width_out, ksize, stride=step, bias=not bn), step)]
            if bn:
                block.append((_NORM[dims](width_out), 1))
            block.append((activation(), 1))
            if dropout > 0:
                block.append((nn.Dropout(p=dropout), None))      # dropout carries no entry in the stride table
            for mod, st in block:
                mods.append(mod)
                if st is not None:
                    per_module_stride.append(st)
            width_in = width_out
            if not final:
                width_out *= unit_scaling
        self.strides = per_module_stride
        self.width = insize_from_outsize(mods, 1)
        self.filled = False
        self.features = nn.Sequential(*mods)
        self.latent_dim = width_in
        self.dims = dims

    def _retarget(self, dense: bool, start: int = 1) -> int:
        """Walk the stack with its stride table: dense=True -> stride 1 / dilation = cumulative stride."""
        cumulative = start
        live = [m for m in self.features.children() if not isinstance(m, nn.Dropout)]   # dropout has no stride entry
        for mod, st in zip(live, self.strides):
            if hasattr(mod, 'dilation'):
                mod.dilation = (cumulative if dense else 1,) * self.dims
            if hasattr(mod, 'stride'):
                mod.stride = (1 if dense else st,) * self.dims
            cumulative *= st
        self.filled = dense
        return cumulative

    def fill(self, stride: int = 1) -> int:
        return self._retarget(True, stride)

    def unfill(self) -> None:
        self._retarget(False)

    def forward(self, x):
        from topaz_b200 import engine
        return engine.features_forward(self, x)


class Conv127(BasicConv):
    def __init__(self, units: int, **kwargs):
        super().__init__([7, 5, 5, 5, 5], units, dims=2, **kwargs)


class Conv63(BasicConv):
    def __init__(self, units: int, **kwargs):
        super().__init__([7, 5, 5, 5], units, dims=2, **kwargs)
