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


def load_pretrained_state(kind: str, name: str):
    return torch.load(pretrained_path(kind, name), map_location='cpu', weights_only=False)


def get_patches(X, patch_size, patch_padding=0, is_3d=False):
    """utils.py:133-168 (including the all-zero-patch skip)."""
    y, x = X.shape[-2:]
    z = X.shape[-3] if is_3d else None
    pad = (patch_padding, patch_padding) * (3 if is_3d else 2)
    X = torch.nn.functional.pad(X, pad)
    y_pad, x_pad = X.shape[-2:]
    z_pad = X.shape[-3] if is_3d else None
    step_size = patch_size - 2 * patch_padding
    patches = []
    for i in range(0, y, step_size):
        for j in range(0, x, step_size):
            i_end = min(i + patch_size, y_pad)
            j_end = min(j + patch_size, x_pad)
            if is_3d:
                for k in range(0, z, step_size):
                    k_end = min(k + patch_size, z_pad)
                    patch = X[..., k:k_end, i:i_end, j:j_end]
                    if patch.abs().sum() == 0:
                        continue
                    patches.append(patch)
            else:
                patch = X[..., i:i_end, j:j_end]
                if patch.abs().sum() == 0:
                    continue
                patches.append(patch)
    return patches


def reconstruct_from_patches(patches, original_shape, patch_size, patch_padding=0, is_3d=False):
    """utils.py:172-193 (float64 result)."""
    y, x = original_shape[-2:]
    z = original_shape[-3] if is_3d else None
    step_size = patch_size - patch_padding * 2
    reassembled = np.zeros(original_shape)
    patch_idx = 0
    for i in range(0, y, step_size):
        for j in range(0, x, step_size):
            if is_3d:
                for k in range(0, z, step_size):
                    patch = patches[patch_idx]
                    reassembled[..., k:k + patch.shape[-3], i:i + patch.shape[-2], j:j + patch.shape[-1]] = patch
                    patch_idx += 1
            else:
                patch = patches[patch_idx]
                reassembled[..., i:i + patch.shape[-2], j:j + patch.shape[-1]] = patch
                patch_idx += 1
    return reassembled


def predict_in_patches(model, X, patch_size, is_3d=False, use_cuda=False):
    '''utils.py:110-130: predict on an image in patches (halo = receptive field // 2) and reassemble.'''
    patch_padding = model.width // 2
    patches = get_patches(X, patch_size, patch_padding=patch_padding, is_3d=is_3d)
    scores = []
    for patch in patches:
        with torch.no_grad():
            patch = patch.cuda() if use_cuda else patch
            score = model(patch).data[0, 0].cpu().numpy()
            score = score[..., patch_padding:-patch_padding, patch_padding:-patch_padding]
            if is_3d:
                score = score[..., patch_padding:-patch_padding, :, :]
        scores.append(score)
    return reconstruct_from_patches(scores, X.shape, patch_size, patch_padding=patch_padding, is_3d=is_3d)
