"""Drop-in mirror of topaz.stats.normalize(method='affine') (reference stats.py:36-46): (x - mean) / std (population
std) as float32 + the metadata dict; statistics and the rescale run on the GPU (tpz_meanstd / tpz_affine).
The GMM normalisation (stats.py:86-214) is outside the B200 hot path."""
import numpy as np
import torch

from topaz_b200 import ops


def normalize(x, alpha=900, beta=1, num_iters=100, sample=1, method='gmm', use_cuda=True, verbose=False):
    if method != 'affine':
        raise NotImplementedError("topaz_b200.stats.normalize: only method='affine' is on the B200 hot path")
    xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()
    stats = ops.meanstd(xd, unbiased=False)
    y = ops.affine(xd, stats)
    mu, std = (float(v) for v in stats.cpu())
    return y.cpu().numpy().astype(np.float32), {'mu': mu, 'std': std, 'pi': 1}
