"""Drop-in for ``topaz.model.classifier.LinearClassifier`` (reference classifier.py:14-66): a feature extractor followed
by a 1x1 convolution to one logit per position.  ``self.classifier`` only stores the (weight, bias) under the reference's
state_dict keys; in the dense forward the 1x1 conv is fused into the epilogue of the last feature convolution."""
import torch.nn as nn

from topaz_b200.model.utils import EngineStateMixin


class LinearClassifier(EngineStateMixin, nn.Module):
    def __init__(self, features, dims=2, patch_size: int = None, padding: int = None, batch_size: int = 1):
        super().__init__()
        head = nn.Conv3d if dims == 3 else nn.Conv2d
        self.features = features
        self.dims = dims
        self.classifier = head(features.latent_dim, 1, 1)
        self.patch_size, self.padding, self.batch_size = patch_size, padding, batch_size

    width = property(lambda self: self.features.width)
    latent_dim = property(lambda self: self.features.latent_dim)

    def fill(self, stride=1):
        """Switch to dense evaluation; returns the cumulative stride of the extractor (4 for ResNet8, 8 for conv63)."""
        return self.features.fill(stride=stride)

    def unfill(self):
        self.features.unfill()

    def forward(self, x):
        from topaz_b200 import engine
        return engine.classifier_forward(self, x)


def classify_patches(classifier, tomo_stack, patch_size=48, padding=36, batch_size=1, volume_num=1, total_volumes=1, verbose=True):
    """3-D evaluation tiling of the reference (classifier.py:69-103); 3-D classifiers are outside the B200 hot path."""
    raise NotImplementedError('topaz_b200: 3-D classifier tiling (classify_patches) is outside the B200 hot path')
