"""Drop-in mirror of the denoising pipeline objects of topaz/denoise.py: Denoise (:245-332), Denoise3D
(:336-377), denoise_image (:382-416).  Normalisation statistics, normalise / de-normalise and the network
all run on the GPU; the only host synchronisation per call is the final device->host copy."""
from __future__ import absolute_import, division, print_function

import sys
from typing import List, Union

import os

import numpy as np
import torch

from topaz_b200 import engine, ops
from topaz_b200.denoising.models import load_model


class Denoise():
    ''' Object for micrograph denoising utilities (reference denoise.py:245-332). '''
    def __init__(self, model: Union[torch.nn.Module, str], use_cuda=True, dims=2):
        if isinstance(model, torch.nn.Module):
            self.model = model
        elif type(model) == str:
            try:
                self.model = load_model(model)
            except NotImplementedError:
                raise
            except Exception:
                raise ValueError('Unable to load model: ' + model)
        else:
            raise TypeError('Unrecognized model:' + str(model))
        if not use_cuda:
            raise RuntimeError('topaz_b200: Denoise requires use_cuda=True; this build has no CPU path')
        self.model = self.model.cuda()
        self.device = next(iter(self.model.parameters())).device
        self.dims = dims
        self.use_cuda = use_cuda

    def __call__(self, input):
        return self._denoise(input)

    @torch.no_grad()
    def _denoise_device(self, input: torch.Tensor) -> torch.Tensor:
        """_denoise without the final host copy: returns the de-normalised prediction on the device."""
        self.model.eval()
        x = input.to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
        stats = ops.meanstd(x, unbiased=True)            # torch .mean() / .std() (denoise.py:283)
        xn = ops.affine(x, stats)                        # (x - mu) / std (denoise.py:284)
        if xn.dim() == self.dims:
            xn = xn[None, None]
        elif xn.dim() == self.dims + 1:
            xn = xn.unsqueeze(1)
        from topaz_b200.denoising.models import DenoiseNet2, AffineDenoise
        if isinstance(self.model, DenoiseNet2):
            pred = engine.fcnn_forward(self.model, xn, denorm_stats=stats)
        elif isinstance(self.model, AffineDenoise):
            pred = engine.affine_forward(self.model, xn, denorm_stats=stats)
        else:
            pred = engine.unet_forward(self.model, xn, denorm_stats=stats)   # pred*std+mu fused (denoise.py:295)
        return pred.squeeze()

    @torch.no_grad()
    def _denoise(self, input: Union[np.ndarray, torch.Tensor]) -> np.ndarray:
        '''Call stored denoising model (reference denoise.py:274-296).'''
        input = torch.from_numpy(input) if type(input) == np.ndarray else input
        return self._denoise_device(input).cpu().numpy()

    def _denoise_crop(self, crop: torch.Tensor) -> torch.Tensor:
        """_denoise_device for one patch crop, replayed from a CUDA graph per crop shape.  A 4096^2 micrograph is 16 crops
        of 4 shapes and ~90 kernel launches each; issued one by one from Python the host becomes the bottleneck (1422
        launches, ~45 ms) although the GPU needs 39 ms.  Graphs are keyed by (shape, parameter versions) and hold their
        own static input / activations / output; TPZ_DENOISE_GRAPH=0 disables them."""
        if os.environ.get('TPZ_DENOISE_GRAPH', '1') == '0':
            return self._denoise_device(crop)
        graphs = self.__dict__.setdefault('_graphs', {})
        key = (tuple(crop.shape), str(crop.device), engine._state_key(self.model))      # includes engine.PRECISION
        hit = graphs.get(key)
        if hit is None:
            if len(graphs) >= 8:                       # weights changed or many shapes: drop the old pools
                graphs.clear()
            try:
                static_in = crop.contiguous().clone()
                cur = torch.cuda.current_stream()
                side = torch.cuda.Stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):          # eager warm-up: builds plans, sets kernel attributes
                    self._denoise_device(static_in)
                cur.wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    static_out = self._denoise_device(static_in)
                # the captured launches hold the ADDRESSES of the packed weights (allocated by the eager warm-up, outside the
                # graph's pool): keep the plans / native handles they belong to alive for as long as the graph can be replayed
                # (engine.PRECISION switched away and back would otherwise replay against freed weights)
                keep = dict(self.model.__dict__.get('_tpz_plans', {}))
                hit = (graph, static_in, static_out, keep)
            except Exception as e:                     # capture not possible: stay eager for this shape
                print(f'topaz_b200: CUDA-graph capture of the denoiser failed ({type(e).__name__}: {e}); running eagerly',
                      file=sys.stderr)
                hit = (None, None, None, None)
            graphs[key] = hit
        graph, static_in, static_out = hit[:3]
        if graph is None:
            return self._denoise_device(crop)
        static_in.copy_(crop)
        graph.replay()
        return static_out

    def _patch_row(self, xd: torch.Tensor, y: torch.Tensor, i: int, patch_size: int, padding: int):
        """All patches whose centres start at row i (reference denoise.py:305-322)."""
        H, W = xd.shape[0], xd.shape[1]
        si, ei = max(0, i - padding), min(H, i + patch_size + padding)
        for j in range(0, W, patch_size):
            sj, ej = max(0, j - padding), min(W, j + patch_size + padding)
            yij = self._denoise_crop(xd[si:ei, sj:ej])
            oi, oj = i - si, j - sj
            y[i:i + patch_size, j:j + patch_size] = yij[oi:oi + patch_size, oj:oj + patch_size]

    @torch.no_grad()
    def denoise_patches_device(self, xd: torch.Tensor, patch_size: int, padding: int = 128) -> torch.Tensor:
        """Patch loop of denoise_patches on a device-resident fp32 micrograph; returns the device result."""
        y = torch.zeros_like(xd)
        for i in range(0, xd.shape[0], patch_size):
            self._patch_row(xd, y, i, patch_size, padding)
        return y

    def _pinned(self, name: str, shape) -> torch.Tensor:
        """Cached pinned staging buffer (host<->device copies run at PCIe speed instead of pageable-memcpy speed)."""
        cache = self.__dict__.setdefault('_pin', {})
        buf = cache.get(name)
        if buf is None or tuple(buf.shape) != tuple(shape):
            buf = torch.empty(tuple(shape), dtype=torch.float32).pin_memory()
            cache[name] = buf
        return buf

    @torch.no_grad()
    def denoise_patches(self, x: Union[np.ndarray, torch.Tensor], patch_size: int, padding: int = 128) -> np.ndarray:
        ''' Denoise 2D micrograph patches (reference denoise.py:299-324).  Patches are cropped, denoised and pasted on the
        device.  For a host image the transfer is pipelined by patch row: the rows a patch row needs are staged through
        pinned memory and uploaded on a copy stream while earlier rows compute, and each finished row is downloaded and
        copied into the result while later rows compute.'''
        x = torch.from_numpy(x) if type(x) == np.ndarray else x
        if x.is_cuda:
            y = self.denoise_patches_device(x.float(), patch_size, padding)
            return y.cpu().numpy()
        x = x.float()
        H, W = x.shape
        stage = self._pinned('in', x.shape)
        # The result lives in a pinned block of torch's caching host allocator and is returned as a numpy view of it
        # (no pageable copy, no first-touch page faults of a fresh 64 MB array); dropping the array recycles the block.
        out = torch.empty((H, W), dtype=torch.float32, pin_memory=True)
        if os.environ.get('TPZ_DENOISE_PIPELINE', '1') == '0':           # A/B switch: whole-image upload / download
            stage.copy_(x)
            y = self.denoise_patches_device(stage.to(self.device, non_blocking=True), patch_size, padding)
            out.copy_(y, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return out.numpy()
        streams = self.__dict__.setdefault('_streams', None) or (torch.cuda.Stream(), torch.cuda.Stream())
        self._streams = streams
        copy_in, copy_out = streams
        main = torch.cuda.current_stream()
        xd = torch.empty((H, W), dtype=torch.float32, device=self.device)
        y = torch.empty((H, W), dtype=torch.float32, device=self.device)
        copy_in.wait_stream(main); copy_out.wait_stream(main)
        # Host staging (pageable numpy -> pinned) of every band runs on a small thread pool from the start: the 64 MB copy
        # costs ~17 ms on one core, which would otherwise sit in front of each band's upload on this thread.
        bands, uploaded = [], 0
        for i in range(0, H, patch_size):
            need = min(H, i + patch_size + padding)
            bands.append((uploaded, need) if need > uploaded else None)
            uploaded = max(uploaded, need)
        pool = self.__dict__.get('_pool')
        if pool is None:
            from concurrent.futures import ThreadPoolExecutor
            pool = self._pool = ThreadPoolExecutor(max_workers=4)

        def stage_rows(lo, hi):
            stage[lo:hi].copy_(x[lo:hi])
        futs = []
        for b in bands:
            if b is None:
                futs.append(None)
                continue
            lo, hi = b
            step = max(1, (hi - lo + 1) // 2)           # two pieces per band: four copies in flight
            futs.append([pool.submit(stage_rows, a, min(hi, a + step)) for a in range(lo, hi, step)])
        last = None
        for bi, i in enumerate(range(0, H, patch_size)):
            if bands[bi] is not None:
                lo, hi = bands[bi]
                for f in futs[bi]:
                    f.result()
                with torch.cuda.stream(copy_in):
                    xd[lo:hi].copy_(stage[lo:hi], non_blocking=True)
                    ev = torch.cuda.Event(); ev.record(copy_in)
                main.wait_event(ev)
            self._patch_row(xd, y, i, patch_size, padding)
            ev = torch.cuda.Event(); ev.record(main)
            hi = min(H, i + patch_size)
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(ev)
                out[i:hi].copy_(y[i:hi], non_blocking=True)      # downloads overlap the following patch rows
                last = torch.cuda.Event(); last.record(copy_out)
        if last is not None:
            last.synchronize()
        return out.numpy()

    @torch.no_grad()
    def denoise(self, x: Union[np.ndarray, torch.Tensor], patch_size=-1, padding=128):
        s = patch_size + padding
        use_patch = (patch_size > 0) and (s < x.shape[0] or s < x.shape[1])
        return self.denoise_patches(x, patch_size, padding=padding) if use_patch else self._denoise(x)

    @torch.no_grad()
    def denoise_device(self, xd: torch.Tensor, patch_size=-1, padding=128) -> torch.Tensor:
        """``denoise`` for a device-resident fp32 micrograph, result left on the device (denoise_image keeps the whole
        normalise -> denoise -> restore chain there)."""
        s = patch_size + padding
        use_patch = (patch_size > 0) and (s < xd.shape[0] or s < xd.shape[1])
        return self.denoise_patches_device(xd, patch_size, padding) if use_patch else self._denoise_device(xd)


class Denoise3D(Denoise):
    ''' Object for denoising tomograms (reference denoise.py:336-377). '''
    def __init__(self, model, use_cuda=True, dims=3):
        super().__init__(model, use_cuda=use_cuda, dims=dims)

    @torch.no_grad()
    def denoise(self, tomo: np.ndarray, patch_size: int = 96, padding: int = 48, batch_size: int = 1,
                volume_num: int = 1, total_volumes: int = 1, verbose: bool = True,
                patch_range=None) -> np.ndarray:
        """``patch_range=(start, stop)`` restricts the work to a slice of the patch list (multi-GPU sharding);
        voxels of other patches stay zero."""
        if patch_size < 1:
            denoised = np.zeros_like(tomo)
            denoised[:] = Denoise._denoise(self, tomo)
            return denoised
        # The whole tomogram goes to the device once (512^3 fp32 = 537 MB); its mean / std (reference: tomo.mean(), tomo.std() on the
        # host, ~0.3 s of numpy per call and per rank) are reduced there with fp64 accumulators and stay there: the per-patch
        # normalise / de-normalise read them as 0-dim tensors, so no host synchronisation happens before the final copy.
        tomo32 = np.ascontiguousarray(tomo, dtype=np.float32)
        td = torch.from_numpy(tomo32).to(self.device)
        stats = ops.meanstd(td, unbiased=False)               # numpy's std is the population std (ddof = 0)
        mu32, std32 = stats[0], stats[1]
        out_d = torch.zeros_like(td)
        pz = [int(np.ceil(n / patch_size)) for n in tomo.shape]
        total = int(np.prod(pz))
        lo, hi = (0, total) if patch_range is None else patch_range
        d = patch_size + 2 * padding
        count = 0
        for p in range(lo, hi):
            i, j, k = (int(v) * patch_size for v in np.unravel_index(p, pz))
            # zero-padded (p+2*pad)^3 crop (reference PatchDataset.__getitem__, datasets.py:426-468)
            x = torch.zeros((d, d, d), dtype=torch.float32, device=self.device)
            si, ei = max(0, i - padding), min(tomo.shape[0], i + patch_size + padding)
            sj, ej = max(0, j - padding), min(tomo.shape[1], j + patch_size + padding)
            sk, ek = max(0, k - padding), min(tomo.shape[2], k + patch_size + padding)
            sic, sjc, skc = padding - i + si, padding - j + sj, padding - k + sk
            x[sic:sic + ei - si, sjc:sjc + ej - sj, skc:skc + ek - sk] = td[si:ei, sj:ej, sk:ek]
            xb = (x[None] - mu32) / std32                     # batch of 1 (DataLoader(batch_size=1))
            y = self._denoise_device(xb) * std32 + mu32       # GPU-bound (12 ms per crop): graph replay measured slower here
            dz, dy, dx = out_d[i:i + patch_size, j:j + patch_size, k:k + patch_size].shape
            out_d[i:i + patch_size, j:j + patch_size, k:k + patch_size] = \
                y[padding:padding + dz, padding:padding + dy, padding:padding + dx]
            count += 1
            if verbose:
                print(f'# [{volume_num}/{total_volumes}] {round(count*100/max(1, hi-lo))}%', file=sys.stderr, end='\r')
        if verbose:
            print(' ' * 100, file=sys.stderr, end='\r')
        # The result comes back through a pinned block of torch's caching host allocator and is returned as a numpy view of it (the 2-D
        # path does the same): no pageable staging copy, no second 512 MB memcpy into a freshly zero-paged array.
        out = torch.empty(tuple(tomo.shape), dtype=torch.float32, pin_memory=out_d.is_cuda)
        out.copy_(out_d, non_blocking=True)
        if out_d.is_cuda:
            torch.cuda.current_stream().synchronize()
        res = out.numpy()
        return res if tomo.dtype == np.float32 else res.astype(tomo.dtype)


def denoise_image(mic: np.ndarray, models: List[Denoise], lowpass=1, cutoff=0, gaus=None, inv_gaus=None,
                  deconvolve=False, deconv_patch=1, patch_size=-1, padding=0, normalize=False, use_cuda=True) -> np.ndarray:
    ''' reference denoise.py:382-416 with the optional lowpass / inverse-Gaussian / deconvolve pre-filters outside the hot path
    (off in every BASELINE config).  The micrograph is uploaded ONCE; the outer normalisation (population mean / std,
    denoise.py:388-389), the optional cutoff and Gaussian pre-filter, every model's (patched) forward, the average over
    models and the final re-normalisation or scale restore (:409-414) all run on the device, and the result comes back in
    one copy -- the reference makes four full-image numpy passes on the host around the network.'''
    if lowpass > 1 or inv_gaus is not None or deconvolve:
        raise NotImplementedError('topaz_b200: lowpass / inverse-Gaussian / deconvolve pre-filters are outside the B200 hot path')
    if not use_cuda:
        raise RuntimeError('topaz_b200: denoise_image requires use_cuda=True; this build has no CPU path')
    dev = models[0].device
    host = torch.from_numpy(np.ascontiguousarray(mic, dtype=np.float32))
    stage = models[0]._pinned('img', host.shape)
    stage.copy_(host)
    with torch.no_grad():
        xd = stage.to(dev, non_blocking=True)
        stats = ops.meanstd(xd, unbiased=False)           # mic.mean(), mic.std() (numpy: population std)
        x = ops.affine(xd, stats)                         # (mic - mu) / std
        if cutoff > 0:
            x = torch.where((x < -cutoff) | (x > cutoff), torch.zeros((), device=dev), x)
        if gaus is not None:                              # topaz_b200.filters.GaussianDenoise (reference denoise.py:397-398)
            x = gaus.forward(x[None, None])[0, 0].contiguous()
        acc = None
        for model in models:
            y = model.denoise_device(x, patch_size=patch_size, padding=padding)
            acc = y.clone() if acc is None and len(models) > 1 else (y if acc is None else acc.add_(y))
        if len(models) > 1:
            acc = acc / len(models)
        if normalize:
            out = ops.affine(acc, ops.meanstd(acc, unbiased=False))          # (mic - mic.mean()) / mic.std()
        else:
            out = ops.affine(acc, stats, inverse=True)                       # std * mic + mu
        res = torch.empty(tuple(out.shape), dtype=torch.float32, pin_memory=True)
        res.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    return res.numpy()
