"""Drop-in mirrors of the U-Net denoisers of topaz/denoising/models.py (UDenoiseNet :74-175,
UDenoiseNetSmall :178-244, UDenoiseNet3D :452-564, load_model :568-625): same constructor arguments and
state_dict keys (enc{i}.0.*, dec{i}.{0,2,4}.*).  nn.Conv* children only hold parameters; forward() runs the
sm_100a kernels (topaz_b200.engine.unet_forward)."""
import sys
from collections import OrderedDict

import torch
from torch import nn

from topaz_b200.model.utils import EngineStateMixin, load_pretrained_state


class _UNetBase(EngineStateMixin, nn.Module):
    _dims = 2

    def _build(self, nf, base_width, top_width, depth):
        conv = nn.Conv3d if self._dims == 3 else nn.Conv2d
        pool = nn.MaxPool3d if self._dims == 3 else nn.MaxPool2d
        act = lambda: nn.LeakyReLU(0.1)
        # encoder: depth conv stages, every one but the last followed by MaxPool(2)
        for i in range(1, depth + 1):
            k = base_width if i == 1 else 3
            mods = [conv(1 if i == 1 else nf, nf, k, padding=k // 2), act()]
            if i < depth:
                mods.append(pool(2))
            setattr(self, f'enc{i}', nn.Sequential(*mods))
        # decoder: level l consumes [upsampled, skip]
        cin = 2 * nf
        for l in range(depth - 1, 1, -1):
            setattr(self, f'dec{l}', nn.Sequential(conv(cin, 2 * nf, 3, padding=1), act(),
                                                   conv(2 * nf, 2 * nf, 3, padding=1), act()))
            cin = 3 * nf
        t = top_width
        self.dec1 = nn.Sequential(conv(2 * nf + 1, 64, t, padding=t // 2), act(),
                                  conv(64, 32, t, padding=t // 2), act(),
                                  conv(32, 1, t, padding=t // 2))

    def forward(self, x):
        from topaz_b200 import engine
        return engine.unet_forward(self, x)


class UDenoiseNet(_UNetBase):
    # U-net from noise2noise paper (reference denoising/models.py:74-175)
    def __init__(self, nf=48, base_width=11, top_width=3):
        super().__init__()
        self._build(nf, base_width, top_width, depth=6)


class UDenoiseNetSmall(_UNetBase):
    # reference denoising/models.py:178-244
    def __init__(self, nf=48, width=11, top_width=3):
        super().__init__()
        self._build(nf, width, top_width, depth=4)


class UDenoiseNet3D(_UNetBase):
    # reference denoising/models.py:452-564
    _dims = 3

    def __init__(self, nf=48, base_width=11, top_width=3):
        super().__init__()
        self._build(nf, base_width, top_width, depth=6)


class DenoiseNet2(EngineStateMixin, nn.Module):
    """`fcnn` denoiser (reference denoising/models.py:52-66): three same-padded width x width convs, LeakyReLU(0.1)."""
    def __init__(self, base_filters, width=11):
        super().__init__()
        self.base_filters = base_filters
        nf = base_filters
        self.net = nn.Sequential(nn.Conv2d(1, nf, width, padding=width // 2), nn.LeakyReLU(0.1),
                                 nn.Conv2d(nf, nf, width, padding=width // 2), nn.LeakyReLU(0.1),
                                 nn.Conv2d(nf, 1, width, padding=width // 2))

    def forward(self, x):
        from topaz_b200 import engine
        return engine.fcnn_forward(self, x)


class AffineDenoise(EngineStateMixin, nn.Module):
    """`affine` denoiser (reference filters.py:40-48): one learned max_size x max_size filter."""
    def __init__(self, max_size=31):
        super().__init__()
        self.filter = nn.Conv2d(1, 1, max_size, padding=max_size // 2)
        self.filter.weight.data.zero_()
        self.filter.bias.data.zero_()

    def forward(self, x):
        from topaz_b200 import engine
        return engine.affine_forward(self, x)


model_name_dict = {
    # 2D models
    'unet': 'unet_L2_v0.2.2.sav',
    'unet-small': 'unet_small_L1_v0.2.2.sav',
    'fcnn': 'fcnn_L1_v0.2.2.sav',
    'affine': 'affine_L1_v0.2.2.sav',
    'unet-v0.2.1': 'unet_L2_v0.2.1.sav',
    # 3D models
    'unet-3d': 'unet-3d-10a-v0.2.4.sav',
    'unet-3d-10a': 'unet-3d-10a-v0.2.4.sav',
    'unet-3d-20a': 'unet-3d-20a-v0.2.4.sav',
}

_ARCH = {
    'unet_L2_v0.2.1.sav': lambda: UDenoiseNet(base_width=7, top_width=3),
    'unet_L2_v0.2.2.sav': lambda: UDenoiseNet(base_width=11, top_width=5),
    'unet_small_L1_v0.2.2.sav': lambda: UDenoiseNetSmall(width=11, top_width=5),
    'fcnn_L1_v0.2.2.sav': lambda: DenoiseNet2(64, width=11),
    'affine_L1_v0.2.2.sav': lambda: AffineDenoise(max_size=31),
    'unet-3d-10a-v0.2.4.sav': lambda: UDenoiseNet3D(base_width=7),
    'unet-3d-20a-v0.2.4.sav': lambda: UDenoiseNet3D(base_width=7),
}


def load_model(name, base_kernel_width=11):
    '''reference denoising/models.py:581-625.'''
    pretrained = name in model_name_dict
    if pretrained:
        name = model_name_dict[name]
    if name in _ARCH:
        model = _ARCH[name]()
    else:
        model = torch.load(name, weights_only=False)
    if pretrained:
        print('# loading pretrained model:', name, file=sys.stderr)
        model.load_state_dict(load_pretrained_state('denoise', name))
    elif type(model) is OrderedDict and '3d' in name:
        state = model
        model = UDenoiseNet3D(base_width=base_kernel_width)
        model.load_state_dict(state)
    model.eval()
    return model


def train_model(*args, **kwargs):
    """Noise2noise denoiser training (reference denoising/models.py:636-758) is outside the B200 hot path."""
    raise NotImplementedError('topaz_b200: denoiser training is outside the B200 hot path; train with the reference and load the .sav')
