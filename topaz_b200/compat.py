"""Make an unmodified Topaz use the B200 modules: ``install()`` aliases the reference's L2 module paths
(`topaz.model.factory`, `topaz.model.classifier`, `topaz.model.features.{resnet,basic}`, `topaz.denoising.models`,
`topaz.methods`) to their topaz_b200 drop-ins in ``sys.modules`` and patches `topaz.algorithms.non_maximum_suppression(_3d)`.  Call it
BEFORE importing `topaz.extract` / `topaz.training` / `topaz.denoise`; the reference's own L3-L5 code (CLI, pipelines)
then runs unchanged on the sm_100a kernels.  Whole-module pickles saved by the reference (`torch.save(model)`,
training.py:601) resolve to the drop-in classes through the same aliases."""
import importlib
import sys

_ALIASES = {
    'topaz.model.factory': 'topaz_b200.model.factory',
    'topaz.model.classifier': 'topaz_b200.model.classifier',
    'topaz.model.features.resnet': 'topaz_b200.model.features.resnet',
    'topaz.model.features.basic': 'topaz_b200.model.features.basic',
    'topaz.denoising.models': 'topaz_b200.denoising.models',
    'topaz.methods': 'topaz_b200.methods',
}
# functions patched INTO reference modules that also hold out-of-scope helpers (match_coordinates, ...)
_FUNCTION_PATCHES = {('topaz.algorithms', 'non_maximum_suppression'): ('topaz_b200.algorithms', 'non_maximum_suppression'),
                     ('topaz.algorithms', 'non_maximum_suppression_3d'): ('topaz_b200.algorithms', 'non_maximum_suppression_3d'),
                     ('topaz.utils.image', 'downsample'): ('topaz_b200.preprocess', 'downsample'),
                     ('topaz.stats', 'normalize'): ('topaz_b200.stats', 'normalize'),
                     ('topaz.stats', 'norm_fit'): ('topaz_b200.stats', 'norm_fit'),
                     ('topaz.stats', 'gmm_fit'): ('topaz_b200.stats', 'gmm_fit'),
                     # training data: crops sampled + augmented on the GPU (falls back to the reference loader without CUDA / in 3-D)
                     ('topaz.training', 'make_data_iterators'): ('topaz_b200.training', 'make_data_iterators')}


# names that reference modules bind with `from X import f` at import time: re-pointed when the consumer module is ALREADY
# loaded (install() after `import topaz.stats` / `import topaz.extract` would otherwise leave them on the reference function)
_REEXPORTS = {('topaz.stats', 'downsample'): ('topaz_b200.preprocess', 'downsample'),
              ('topaz.extract', 'non_maximum_suppression'): ('topaz_b200.algorithms', 'non_maximum_suppression'),
              ('topaz.extract', 'non_maximum_suppression_3d'): ('topaz_b200.algorithms', 'non_maximum_suppression_3d')}


def install(names=None):
    """Alias the listed reference module paths (default: all) to the topaz_b200 implementations."""
    for ref, ours in _ALIASES.items():
        if names is not None and ref not in names:
            continue
        mod = importlib.import_module(ours)
        sys.modules[ref] = mod
        parent, _, leaf = ref.rpartition('.')
        try:                                      # `import topaz.model.classifier as C` (training.py:15) resolves the leaf as
            pkg = importlib.import_module(parent)  # an ATTRIBUTE of the (real, empty) parent package: import it and bind
        except ImportError:
            pkg = sys.modules.get(parent)
        if pkg is not None:
            setattr(pkg, leaf, mod)
    for (ref_mod, fn), (our_mod, our_fn) in _FUNCTION_PATCHES.items():
        if names is not None and ref_mod not in names:
            continue
        try:
            target = importlib.import_module(ref_mod)
        except ImportError:
            continue
        if not hasattr(target, '_tpz_orig_' + fn):
            setattr(target, '_tpz_orig_' + fn, getattr(target, fn))
        setattr(target, fn, getattr(importlib.import_module(our_mod), our_fn))
    for (ref_mod, fn), (our_mod, our_fn) in _REEXPORTS.items():
        target = sys.modules.get(ref_mod)
        if target is not None and (names is None or ref_mod in names) and hasattr(target, fn):
            setattr(target, fn, getattr(importlib.import_module(our_mod), our_fn))
    return sorted(_ALIASES if names is None else names)


def uninstall():
    for ref in _ALIASES:
        sys.modules.pop(ref, None)
