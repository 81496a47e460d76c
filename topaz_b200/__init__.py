"""topaz_b200 — B200-native (sm_100a) implementation of the Topaz dense-CNN hot path.

Public surface mirrors the reference modules it replaces:
  topaz_b200.model.factory.load_model / get_feature_extractor, topaz_b200.model.classifier.LinearClassifier,
  topaz_b200.model.features.{resnet,basic}, topaz_b200.extract.score_images,
  topaz_b200.denoising.models.{UDenoiseNet,UDenoiseNet3D,load_model}, topaz_b200.denoise.{Denoise,Denoise3D},
  topaz_b200.methods.GE_binomial.
All arithmetic runs in the CUDA kernels of libtopaz_b200.so (include/topaz_b200.h); there is no CPU path.
"""
__version__ = '0.1.0'
