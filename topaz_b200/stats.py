"""GPU drop-in for topaz.stats.normalize / norm_fit / gmm_fit (reference stats.py:36-214).

method='affine' (stats.py:38-46): (x - mean) / std (population std) as float32 + metadata, statistics and rescale on the
GPU (tpz_meanstd / tpz_affine).

method='gmm' (stats.py:49-214): 12 initialisations of a shared-variance 2-component Gaussian mixture with a Beta(alpha,
beta) prior on the mixing weight, EM until the log-posterior improves by <= 1e-3 or `num_iters`; the image is scaled by
the brighter component of the best fit.  The 11 mixture fits advance in lockstep: every EM iteration is ONE fused pass over the pixels for all of
them (tpz_gmm_sums: responsibilities, log-likelihood and all sufficient statistics, fp64 accumulation); the M step (a handful of scalars)
runs on the host in float64.  The quantile initialisation (np.quantile, stats.py:91) uses exact order statistics from
a 3-pass radix select (tpz_select_hist).  The reference runs the same arithmetic in float32 tensors, so its stopping
iteration can differ by rounding noise; results agree to ~1e-4 relative (tests)."""
import math

import numpy as np
import torch

from topaz_b200 import ops

_PIS = (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 0.95, 0.98, 1)


def _xlogy(c, v):
    return 0.0 if c == 0 else (c * math.log(v) if v > 0 else -math.inf * c)


def _beta_logpdf(p, a, b):
    """log Beta(p; a, b) (scipy.stats.beta.logpdf, stats.py:167,203)."""
    return _xlogy(a - 1, p) + _xlogy(b - 1, 1 - p) + math.lgamma(a + b) - math.lgamma(a) - math.lgamma(b)


def _beta_pdf_at_one(a, b):
    if b == 1:
        return float(a)
    return 0.0 if b > 1 else math.inf


def _key_to_float(key: int) -> float:
    bits = (key & 0x7FFFFFFF) if key & 0x80000000 else (~key & 0xFFFFFFFF)
    return float(np.array([bits], dtype=np.uint32).view(np.float32)[0])


def order_statistics(xd: torch.Tensor, ranks) -> np.ndarray:
    """Exact k-th smallest values (0-based ranks) of a device fp32 tensor: three histogram passes over the data."""
    ranks = [int(r) for r in ranks]
    c0 = np.cumsum(ops.select_hist(xd, 0))
    lvl1 = {}                                     # rank -> (top12 bin, residual rank)
    for r in ranks:
        b = int(np.searchsorted(c0, r, side='right'))
        lvl1[r] = (b, r - (int(c0[b - 1]) if b else 0))
    p1 = sorted({b for b, _ in lvl1.values()})
    h1 = ops.select_hist(xd, 1, p1)
    lvl2 = {}
    for r, (b, rr) in lvl1.items():
        c = np.cumsum(h1[p1.index(b)])
        bb = int(np.searchsorted(c, rr, side='right'))
        lvl2[r] = ((b << 12) | bb, rr - (int(c[bb - 1]) if bb else 0))
    p2 = sorted({b for b, _ in lvl2.values()})
    out = {}
    for lo in range(0, len(p2), 64):
        chunk = p2[lo:lo + 64]
        h2 = ops.select_hist(xd, 2, chunk)
        for r, (b, rr) in lvl2.items():
            if b in chunk:
                c = np.cumsum(h2[chunk.index(b)])
                out[r] = _key_to_float((b << 8) | int(np.searchsorted(c, rr, side='right')))
    return np.array([out[r] for r in ranks], dtype=np.float64)


def quantiles(xd: torch.Tensor, qs) -> np.ndarray:
    """np.quantile(x, qs) (default linear interpolation) from exact order statistics."""
    n = xd.numel()
    pos = np.asarray(qs, dtype=np.float64) * (n - 1)
    lo = np.floor(pos).astype(np.int64)
    hi = np.minimum(lo + 1, n - 1)
    ranks = sorted(set(lo.tolist()) | set(hi.tolist()))
    vals = dict(zip(ranks, order_statistics(xd, ranks)))
    a = np.array([vals[int(i)] for i in lo]); b = np.array([vals[int(i)] for i in hi])
    t = pos - lo
    return np.where(t >= 0.5, b - (b - a) * (1 - t), a + (b - a) * t)


class _Image:
    """Flat device pixels + the scalars every fit shares (shifted coordinates xc = x - shift keep the fp64 sums
    well conditioned for raw micrographs whose mean is far above their contrast)."""

    def __init__(self, xd: torch.Tensor):
        self.x = xd.contiguous().view(-1)
        self.n = self.x.numel()
        self.work = torch.empty(12 * 7, dtype=torch.float64, device=xd.device)
        st = ops.meanstd(self.x, unbiased=True).cpu().numpy().astype(np.float64)
        self.mean, self.var_unbiased = float(st[0]), float(st[1]) ** 2
        self.shift = self.mean

    def sums(self, sets8):
        return ops.gmm_sums(self.x, self.shift, sets8, self.work)


class _Fit:
    """One EM run (stats.py:122-214, share_var=True): host-side state machine around the device sums."""

    def __init__(self, img: _Image, pi, split, alpha, beta, scale, tol, num_iters):
        self.img, self.pi, self.split = img, float(pi), float(split)
        self.alpha, self.beta, self.scale, self.tol, self.num_iters = alpha, beta, scale, tol, num_iters
        self.stage, self.it, self.done = 'init', 0, False
        self.mu0 = self.mu1 = self.var = 0.0
        self.logp = self.logp_cur = None

    def request(self):
        """Parameter row for tpz_gmm_sums: {mode, split, mu0-shift, mu1-shift, var0, var1, log(1-pi), log(pi)}."""
        if self.stage == 'init':
            return (0.0, self.split, 0.0, 0.0, 1.0, 1.0, 0.0, 0.0)
        sh = self.img.shift
        lp0 = math.log1p(-self.pi) if self.pi < 1 else -math.inf
        lp1 = math.log(self.pi) if self.pi > 0 else -math.inf
        return (1.0, 0.0, self.mu0 - sh, self.mu1 - sh, self.var, self.var, lp0, lp1)

    def _m_step(self, s):
        """stats.py:138-153 / 176-192 from the sufficient statistics: component means, shared variance."""
        img = self.img
        _, S0, S1, Sx0, Sx1, Sxx0, Sxx1 = s
        mu0 = Sx0 / S0 if S0 > 0 else img.mean - img.shift
        mu1 = Sx1 / S1 if S1 > 0 else img.mean - img.shift
        self.var = ((Sxx0 - 2 * mu0 * Sx0 + mu0 * mu0 * S0) + (Sxx1 - 2 * mu1 * Sx1 + mu1 * mu1 * S1)) / img.n
        self.mu0, self.mu1 = mu0 + img.shift, mu1 + img.shift

    def consume(self, s):
        if self.stage == 'init':                              # hard-split statistics -> first parameters
            self._m_step(s)
            self.stage = 'first'
            return
        # `s` was computed under the current parameters: their log-posterior, and the statistics for the next M step.
        # The reference compares float32 tensors (stats.py:167,203,209), hence the rounding.
        logp = np.float32(self.scale * s[0] + _beta_logpdf(self.pi, self.alpha, self.beta))
        if self.stage == 'first':
            self.logp = self.logp_cur = logp
            self.stage = 'em'
        else:
            self.logp = logp
            if logp - self.logp_cur <= self.tol or self.it >= self.num_iters:
                self.done = True
                return
            self.logp_cur = logp
        if self.num_iters < 1:
            self.done = True
            return
        self.it += 1
        a = self.alpha + s[2]
        b = self.beta + self.img.n - s[2]
        self.pi = (a - 1) / (a + b - 2)                       # MAP estimate under the Beta prior (stats.py:176-179)
        self._m_step(s)

    def result(self):
        return float(self.logp), self.mu0, self.var, self.mu1, self.var, self.pi


def _run_fits(img: _Image, fits):
    """Advance all unfinished fits in lockstep: one device pass per EM iteration for the whole group."""
    while True:
        active = [f for f in fits if not f.done]
        if not active:
            return
        sums = img.sums([f.request() for f in active])
        for f, s in zip(active, sums):
            f.consume(s)


def gmm_fit(x, pi=0.5, split=None, alpha=0.5, beta=0.5, scale=1, tol=1e-3, num_iters=100, share_var=True, verbose=False):
    """stats.py:122-214 -> (logp, mu0, var0, mu1, var1, pi) as Python floats.  share_var=False is not on this path."""
    if not share_var:
        raise NotImplementedError('topaz_b200.stats.gmm_fit: share_var=False is not implemented')
    img = _Image(_to_device(x))
    if split is None:
        split = float(quantiles(img.x, [1 - pi])[0])
    fit = _Fit(img, pi, split, alpha, beta, scale, tol, num_iters)
    _run_fits(img, [fit])
    return fit.result()


def norm_fit(x, alpha=900, beta=1, scale=1, num_iters=100, use_cuda=True, verbose=False):
    """stats.py:86-119 -> (mu, std, pi, logp, mus, stds, pis, logps)."""
    img = _Image(_to_device(x))
    pis = np.array(_PIS, dtype=np.float64)
    splits = quantiles(img.x, 1 - pis)
    fits = {i: _Fit(img, pis[i], splits[i], alpha, beta, scale, 1e-3, num_iters) for i in range(len(pis)) if pis[i] != 1}
    _run_fits(img, list(fits.values()))
    logps, mus, stds = np.zeros(len(pis)), np.zeros(len(pis)), np.zeros(len(pis))
    for i in range(len(pis)):
        if i in fits:
            logp, _, _, mu, var, pi = fits[i].result()
        else:                                                                # single component (stats.py:103-106)
            mu, var, pi = img.mean, img.var_unbiased, 1.0
            logp = float(np.float32(scale * (-(img.n - 1) / 2.0 - 0.5 * img.n * math.log(2 * math.pi * var)) + _beta_pdf_at_one(alpha, beta)))
        pis[i], logps[i], mus[i], stds[i] = pi, logp, mu, math.sqrt(var)
    i = int(np.argmax(logps))
    return mus[i], stds[i], pis[i], logps[i], mus, stds, pis, logps


def _to_device(x) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return ops.to_device(x.float().contiguous())
    return ops.to_device(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)))


def normalize_device(xd: torch.Tensor, alpha=900, beta=1, num_iters=100, sample=1, method='gmm'):
    """Device fp32 image -> (device fp32 normalised image, metadata)."""
    ops.require_cuda(xd, 'image')
    xd = xd.contiguous().float()
    if method == 'affine':
        stats = ops.meanstd(xd, unbiased=False)
        mu, std = (float(v) for v in stats.cpu())
        return ops.affine(xd, stats), {'mu': mu, 'std': std, 'pi': 1}
    x_sample, scale = xd, 1
    if sample > 1:                                                           # stats.py:53-58 (host RNG, numpy global state)
        n = int(np.round(xd.numel() / sample))
        scale = xd.numel() / n
        idx = np.random.choice(xd.numel(), size=n, replace=False)
        x_sample = xd.view(-1)[torch.from_numpy(idx).to(xd.device)]
    mu, std, pi, logp, mus, stds, pis, logps = norm_fit(x_sample, alpha=alpha, beta=beta, scale=scale, num_iters=num_iters)
    stats = torch.tensor([mu, std], dtype=torch.float32, device=xd.device)
    meta = {'mu': mu, 'std': std, 'pi': pi, 'logp': logp, 'mus': mus, 'stds': stds, 'pis': pis, 'logps': logps,
            'alpha': alpha, 'beta': beta, 'sample': sample}
    return ops.affine(xd, stats), meta


def normalize(x, alpha=900, beta=1, num_iters=100, sample=1, method='gmm', use_cuda=True, verbose=False):
    y, meta = normalize_device(_to_device(x), alpha, beta, num_iters, sample, method)
    return y.cpu().numpy().astype(np.float32), meta
