"""GPU drop-in for topaz.utils.image.downsample (reference utils/image.py:38-61): Fourier-crop downsampling.

The reference computes rfft2 -> keep the low-frequency block -> irfft2.  That map is linear and separable, so it is
applied here as two dense products on the tensor cores (3xTF32, fp32-level accuracy) with real operators built once per
(input shape, output shape):

    f = s * ( Rr @ x @ Ca^T  +  Ri @ x @ Cb^T )

Rr + i Ri [m x M] is ifft_m o select o fft_M along rows, where select keeps input frequencies -ceil(m/2) .. m//2-1
(`F[0:m//2]` and `F[-m//2:]`, image.py:49-51 -- note the set is not symmetric, which is what makes Ri non-zero);
Ca / Cb [n x N] are the cosine / sine parts of irfft_n o crop(0 .. n//2) o rfft_N along columns, including numpy's
C2R convention of ignoring the imaginary part of the DC and Nyquist terms; s = (m n)/(M N) (image.py:54-56)."""
from functools import lru_cache

import numpy as np
import torch

from topaz_b200 import ops


def _ratio(num_angle_num, num_angle_den, den_angle_num, den_angle_den, count):
    """sin(pi*a/b) / sin(pi*c/d) with integer a, c (exact argument reduction); where the denominator vanishes the limit
    count*cos(.)/cos(.) is used."""
    a = np.mod(num_angle_num, 2 * num_angle_den).astype(np.float64)
    c = np.mod(den_angle_num, 2 * den_angle_den).astype(np.float64)
    sn, sd = np.sin(np.pi * a / num_angle_den), np.sin(np.pi * c / den_angle_den)
    cn, cd = np.cos(np.pi * a / num_angle_den), np.cos(np.pi * c / den_angle_den)
    zero = np.mod(den_angle_num, den_angle_den) == 0
    return np.where(zero, count * cn / np.where(zero, cd, 1.0), sn / np.where(zero, 1.0, sd))


@lru_cache(maxsize=8)
def _operators(M: int, N: int, m: int, n: int, device: str):
    """Device operators: RT = [Rr | Ri] as [m x 2*Mp] and CT = (Ca^T, Cb^T) as two [Np x np_] blocks, zero padded so that
    K dimensions are multiples of 16 and the output width a multiple of 32."""
    Mp, Np, np_ = -(-M // 16) * 16, -(-N // 16) * 16, -(-n // 32) * 32
    # rows: (1/m) sum_{f=a}^{b} exp(i alpha f), alpha = 2 pi (j/m - k/M) = 2 pi t/(m M), t = j M - k m
    j = np.arange(m, dtype=np.int64)[:, None]
    k = np.arange(M, dtype=np.int64)[None, :]
    t = j * M - k * m
    a, b = -((m + 1) // 2), m // 2 - 1
    dirichlet = _ratio(t, M, t, m * M, m)                      # sin(m alpha/2) / sin(alpha/2)
    tc = np.mod((a + b) * t, 2 * m * M).astype(np.float64)     # phase exp(i alpha (a+b)/2) = exp(i pi (a+b) t/(m M))
    scale = (m * n) / (M * N) / m
    RT = np.zeros((m, 2 * Mp), dtype=np.float32)
    RT[:, :M] = scale * dirichlet * np.cos(np.pi * tc / (m * M))
    RT[:, Mp:Mp + M] = scale * dirichlet * np.sin(np.pi * tc / (m * M))
    # columns: sum_c (w_c/n) exp(i c beta), beta = 2 pi (q/N - l/n) = 2 pi u/(n N), u = q n - l N
    q = np.arange(N, dtype=np.int64)[:, None]
    l = np.arange(n, dtype=np.int64)[None, :]
    u = q * n - l * N
    half = np.mod(u, 2 * n * N).astype(np.float64) * (np.pi / (n * N))      # beta / 2
    zero = np.mod(u, n * N) == 0
    sin_half = np.where(zero, 1.0, np.sin(half))
    if n % 2:
        re = _ratio(u, N, u, n * N, n)                                          # sin(n beta/2)/sin(beta/2)
        im = np.where(zero, 0.0, (np.cos(half) - np.cos(np.mod(u, 2 * N) * (np.pi / N))) / sin_half)
    else:
        um = (n - 1) * u
        nb2 = np.mod(u, 2 * N) * (np.pi / N)                                     # n beta / 2
        re = _ratio(um, n * N, u, n * N, n - 1) + np.cos(nb2)
        im = np.where(zero, 0.0, (np.cos(half) - np.cos(np.mod(um, 2 * n * N) * (np.pi / (n * N)))) / sin_half) + np.sin(nb2)
    CaT = np.zeros((Np, np_), dtype=np.float32)
    CbT = np.zeros((Np, np_), dtype=np.float32)
    CaT[:N, :n] = re / n
    CbT[:N, :n] = im / n
    dev = torch.device(device)
    return torch.from_numpy(RT).to(dev), torch.from_numpy(CaT).to(dev), torch.from_numpy(CbT).to(dev), Mp, Np, np_


def downsample_device(xd: torch.Tensor, factor=1, shape=None) -> torch.Tensor:
    """Device fp32 [M, N] -> device fp32 [m, n]."""
    ops.require_cuda(xd, 'image')
    M, N = xd.shape
    if shape is None:
        shape = (int(M / factor), int(N / factor))
    m, n = shape
    RT, CaT, CbT, Mp, Np, np_ = _operators(M, N, m, n, str(xd.device))
    # The map preserves constants and is linear, so it is applied to the standardised image (x - mean)/std and undone
    # afterwards: raw micrographs carry a mean far above their contrast, and the tensor cores' truncating fp32
    # accumulation would otherwise add a systematic offset proportional to that mean.
    xd = xd.contiguous().float()
    stats = ops.meanstd(xd, unbiased=False)
    stats = torch.stack([stats[0], stats[1].clamp_min(1e-30)])           # constant image: (x - mean)/tiny = 0
    xn = ops.affine(xd, stats)
    if Np != N:
        xp = torch.zeros((M, Np), dtype=torch.float32, device=xd.device)
        xp[:, :N] = xn
    else:
        xp = xn
    T = torch.zeros((2 * Mp, np_), dtype=torch.float32, device=xd.device)        # [x Ca^T ; x Cb^T], K-padding rows stay 0
    ops.gemm_f32(xp, CaT, T[:M])
    ops.gemm_f32(xp, CbT, T[Mp:Mp + M])
    out = torch.empty((m, np_), dtype=torch.float32, device=xd.device)
    ops.gemm_f32(RT, T, out)
    return ops.affine(out[:, :n].contiguous(), stats, inverse=True)


def downsample(x, factor=1, shape=None):
    """numpy in -> numpy out with the reference's signature (utils/image.py:38); leading batch dims are looped."""
    x = np.asarray(x)
    if x.ndim > 2:
        return np.stack([downsample(xi, factor, shape) for xi in x])
    xd = ops.to_device(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)))
    return downsample_device(xd, factor, shape).cpu().numpy().astype(x.dtype if x.dtype.kind == 'f' else np.float32)
