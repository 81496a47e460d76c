"""Drop-in mirror of topaz/model/classifier.py:14-66 (LinearClassifier = features + 1x1 conv to one logit)."""
from __future__ import division, print_function

import torch
import torch.nn as nn


class LinearClassifier(nn.Module):
    '''A simple convolutional layer without non-linear activation.'''

    def __init__(self, features, dims=2, patch_size: int = None, padding: int = None, batch_size: int = 1):
        super().__init__()
        self.features = features
        self.dims = dims
        conv = nn.Conv3d if dims == 3 else nn.Conv2d
        self.classifier = conv(features.latent_dim, 1, 1)   # parameter container
        self.patch_size = patch_size
        self.padding = padding
        self.batch_size = batch_size

    @property
    def width(self):
        return self.features.width

    @property
    def latent_dim(self):
        return self.features.latent_dim

    def fill(self, stride=1):
        return self.features.fill(stride=stride)

    def unfill(self):
        self.features.unfill()

    def forward(self, x):
        from topaz_b200 import engine
        return engine.classifier_forward(self, x)
