"""Drop-in mirror of topaz/model/factory.py:15-64."""
from __future__ import print_function, division

import torch

from topaz_b200.model.features.basic import BasicConv
from topaz_b200.model.features.resnet import ResNet16, ResNet8
from topaz_b200.model.classifier import LinearClassifier
from topaz_b200.model.utils import load_pretrained_state

resnet16 = ResNet16
resnet8 = ResNet8


def conv127(*args, **kwargs):
    return BasicConv([7, 5, 5, 5, 5], *args, **kwargs)


def conv63(*args, **kwargs):
    return BasicConv([7, 5, 5, 5], *args, **kwargs)


def conv31(*args, **kwargs):
    return BasicConv([7, 5, 5], *args, **kwargs)


_CTORS = {'resnet16': resnet16, 'resnet8': resnet8, 'conv127': conv127, 'conv63': conv63, 'conv31': conv31}


def get_feature_extractor(model, *args, **kwargs):
    if model not in _CTORS:
        raise ValueError(f'unknown feature extractor {model!r}')
    return _CTORS[model](*args, **kwargs)


_PRETRAINED = {
    'resnet16': ('resnet16_u64.sav', ResNet16, 64), 'resnet16_u64': ('resnet16_u64.sav', ResNet16, 64),
    'resnet16_u32': ('resnet16_u32.sav', ResNet16, 32),
    'resnet8': ('resnet8_u64.sav', ResNet8, 64), 'resnet8_u64': ('resnet8_u64.sav', ResNet8, 64),
    'resnet8_u32': ('resnet8_u32.sav', ResNet8, 32),
}


def load_model(path):
    if path in _PRETRAINED:
        name, ctor, units = _PRETRAINED[path]
        model = LinearClassifier(ctor(units=units, bn=False))
        model.load_state_dict(load_pretrained_state('detector', name))
        return model
    return torch.load(path, weights_only=False)
