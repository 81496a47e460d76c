"""Mirror of the parts of topaz/model/utils.py on the hot path: insize_from_outsize (:39-68),
predict_in_patches / get_patches / reconstruct_from_patches (:110-193)."""
from __future__ import division, print_function

import os
import numpy as np
import torch


def insize_from_outsize(layers, outsize):
    """calculates in input size of a convolution stack given the layers and output size (utils.py:39-68)"""
    for layer in layers[::-1]:
        def attr(name, default):
            v = getattr(layer, name, default)
            return v[0] if type(v) is tuple else v
        kernel_size, stride = attr('kernel_size', 1), attr('stride', 1)
        pad, dilation = attr('padding', 0), attr('dilation', 1)
        outsize = (outsize - 1) * stride + 1 + (kernel_size - 1) * dilation - 2 * pad
    return outsize


def pretrained_path(kind: str, name: str) -> str:
    """Locate a packaged Topaz weight file (topaz/pretrained/<kind>/<name>).

    Search order: $TOPAZ_PRETRAINED_DIR/<kind>/<name>, an installed ``topaz`` package, and the
    reference checkout used in the build container."""
    cands = []
    env = os.environ.get('TOPAZ_PRETRAINED_DIR')
    if env:
        cands.append(os.path.join(env, kind, name))
    try:
        import importlib.util
        spec = importlib.util.find_spec('topaz')
        if spec is not None and spec.submodule_search_locations:
            cands.append(os.path.join(list(spec.submodule_search_locations)[0], 'pretrained', kind, name))
    except Exception:
        pass
    cands.append(os.path.join('/root/reference/topaz/pretrained', kind, name))
    for c in cands:
        if os.path.exists(c):
            return c
    raise RuntimeError(f'Could not locate pretrained weights {kind}/{name}; set TOPAZ_PRETRAINED_DIR')


def load_state_dict_from_pkg(pkg: str, path: str, map_location='cpu'):
    """Signature-compatible with topaz.model.utils.load_state_dict_from_pkg (reference utils.py:12-36): ``path`` is
    'pretrained/<kind>/<file>' (possibly with a leading '../')."""
    parts = path.replace('\\', '/').split('/')
    return load_pretrained_state(parts[-2], parts[-1])


def load_pretrained_state(kind: str, name: str):
    return torch.load(pretrained_path(kind, name), map_location='cpu', weights_only=False)


def _patch_origins(shape, step, is_3d):
    """Origins of the patches in the reference's iteration order: y outer, x middle, z inner (utils.py:146-166, 180-191)."""
    import itertools
    y, x = shape[-2:]
    ranges = [range(0, y, step), range(0, x, step)]
    if is_3d:
        ranges.append(range(0, shape[-3], step))
    return itertools.product(*ranges)


def get_patches(X, patch_size, patch_padding=0, is_3d=False):
    """Split a padded image/volume into overlapping patches (reference utils.py:133-168): the input is zero-padded by
    ``patch_padding`` on every side, patches of ``patch_size`` start every ``patch_size - 2*patch_padding`` pixels and are
    clipped at the padded border; all-zero patches are dropped (as the reference does)."""
    ndim_sp = 3 if is_3d else 2
    Xp = torch.nn.functional.pad(X, (patch_padding, patch_padding) * ndim_sp)
    step = patch_size - 2 * patch_padding
    out = []
    for org in _patch_origins(X.shape, step, is_3d):
        i, j = org[0], org[1]
        window = (slice(i, min(i + patch_size, Xp.shape[-2])), slice(j, min(j + patch_size, Xp.shape[-1])))
        if is_3d:
            k = org[2]
            window = (slice(k, min(k + patch_size, Xp.shape[-3])),) + window
        patch = Xp[(Ellipsis,) + window]
        if patch.abs().sum() == 0:
            continue
        out.append(patch)
    return out


def reconstruct_from_patches(patches, original_shape, patch_size, patch_padding=0, is_3d=False):
    """Paste the halo-cropped patch scores back (reference utils.py:172-193); the result is float64 like the reference."""
    step = patch_size - 2 * patch_padding
    canvas = np.zeros(original_shape)
    for patch, org in zip(patches, _patch_origins(original_shape, step, is_3d)):
        i, j = org[0], org[1]
        window = (slice(i, i + patch.shape[-2]), slice(j, j + patch.shape[-1]))
        if is_3d:
            window = (slice(org[2], org[2] + patch.shape[-3]),) + window
        canvas[(Ellipsis,) + window] = patch
    return canvas


def predict_in_patches(model, X, patch_size, is_3d=False, use_cuda=False):
    """Score an image patch-wise with a halo of half the receptive field and stitch (reference utils.py:110-130)."""
    halo = model.width // 2
    crop = (Ellipsis,) + ((slice(halo, -halo),) * 3 if is_3d else (slice(halo, -halo),) * 2)
    scores = []
    with torch.no_grad():
        for patch in get_patches(X, patch_size, patch_padding=halo, is_3d=is_3d):
            dev = patch.cuda() if use_cuda else patch
            scores.append(model(dev).data[0, 0].cpu().numpy()[crop])
    return reconstruct_from_patches(scores, X.shape, patch_size, patch_padding=halo, is_3d=is_3d)


class EngineStateMixin:
    """The engine keeps per-module caches (packed fp16 weights and launch plans, the flat training buffers, the activation
    tape) in the instance ``__dict__`` under ``_tpz_*`` names.  They are rebuilt on demand and must not travel with pickles
    -- the reference saves WHOLE modules (``torch.save(classifier, path)``, training.py:600-601) -- or deep copies."""

    def __getstate__(self):
        return {k: v for k, v in self.__dict__.items() if not k.startswith('_tpz_')}
