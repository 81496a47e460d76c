"""Secondary BASELINE.json workloads (configs 3-5) and the same-GPU library baseline, as functions that bench.py folds into
the `extra` block of its JSON line (and tools/bench_extra.py prints one by one).  Every entry carries its own clock sample.

  cfg3  UDenoiseNet `unet` through Denoise.denoise(x, patch_size=1024, padding=500) on 4096x4096 raw-like micrographs,
        image-sharded over ranks (weak scaling, no collective)
  cfg4  GE_binomial.step, resnet8_u32, global minibatch of 256 crops of 71x71 sharded over ranks, NCCL gradient all-reduce
        (strong scaling); the collectives are also timed in isolation
  cfg5  UDenoiseNet3D `unet-3d-10a` through Denoise3D.denoise(tomo, patch_size=96, padding=48) on an S^3 tomogram, the patch
        list sharded over ranks (strong scaling, no collective)
  gpu_library_baseline  the reference's op sequences on torch + cuDNN (TF32 convolutions, the reference's own GPU path) on the
        same GPU in the same process (tools/torch_cudnn_baseline.py builds them; random weights, timing only)
No oracle import here: the oracle is the checker, not a timed path."""
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tools')):
    if p not in sys.path:
        sys.path.insert(0, p)


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons during a timed region.  Uses NVML in-process (initialised before the timed
    region): spawning `nvidia-smi` every 200 ms re-initialises NVML each time, which takes a driver-wide lock and stalls
    the host-side CUDA calls of the end-to-end leg by tens of ms.  Falls back to nvidia-smi if pynvml is unavailable."""
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []      # rows: (sm_mhz, sm_max_mhz, [active reason flags])
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: address the GPU by the UUID torch reports
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            uuid = uuid if uuid.startswith('GPU-') else 'GPU-' + uuid
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid)
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        flags = [bool(r & n.nvmlClocksThrottleReasonHwSlowdown), bool(r & n.nvmlClocksThrottleReasonHwThermalSlowdown),
                 bool(r & n.nvmlClocksThrottleReasonSwThermalSlowdown), bool(r & n.nvmlClocksThrottleReasonSwPowerCap)]
        self.rows.append((int(sm), int(mx), flags))

    def _sample_smi(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}', '--format=csv,noheader,nounits'],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            c = [v.strip() for v in out.split(',')]
            if c[0].isdigit():
                self.rows.append((int(c[0]), int(c[1]) if c[1].isdigit() else None, [v.lower().startswith('active') for v in c[2:6]]))

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.05 if self.nvml is not None else 0.2)

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2)
        return self.summary()

    def summary(self):
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unavailable'])
        sm = sorted(r[0] for r in self.rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[2][i] for r in self.rows)]
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=self.rows[0][1], reasons=reasons, samples=len(self.rows),
                    source='nvml' if self.nvml is not None else 'nvidia-smi')


class Ctx:
    """rank / world / device plumbing shared by the workloads (one process per GPU; NCCL initialised by the caller)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank = int(os.environ.get('RANK', 0))
        self.world = int(os.environ.get('WORLD_SIZE', 1))
        self.local = int(os.environ.get('LOCAL_RANK', 0))

    def sync(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def maxval(self, v):
        t = torch.tensor([v], dtype=torch.float64, device='cuda')
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def _load(model, sd):
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return model


def cfg3_denoise2d(ctx: Ctx, steps: int = 4, size: int = 4096):
    from common import gold, weights_of
    from topaz_b200 import ops
    from topaz_b200.denoising.models import UDenoiseNet
    from topaz_b200.denoise import Denoise
    dn = Denoise(_load(UDenoiseNet(base_width=11, top_width=5), weights_of(gold('unet_pretrained'))))
    imgs = [(10 + 3 * np.random.default_rng(3000 + ctx.rank * 100 + i).standard_normal((size, size))).astype(np.float32) for i in range(2)]
    y = None
    for _ in range(3):                                  # warm-up: plans, CUDA graphs per crop shape, pinned staging; the result is
        y = dn.denoise(imgs[0], patch_size=1024, padding=500)   # held like in the timed loop, so the pinned-host allocator owns both
                                                                # result blocks the steady state alternates between (a 64 MB
                                                                # cudaHostAlloc inside the timed region costs ~20 ms)
    xd = torch.from_numpy(imgs[0]).cuda()
    for _ in range(2):
        dn.denoise_patches_device(xd, 1024, 500)
    ctx.sync()
    clk = ClockSampler(ctx.local); clk.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ops.LAUNCH_COUNT
    e0.record()
    for i in range(steps):
        dn.denoise_patches_device(xd, 1024, 500)
    e1.record(); ctx.sync()
    dev_ms = ctx.maxval(e0.elapsed_time(e1))
    launches = (ops.LAUNCH_COUNT - l0) / steps
    t0 = time.perf_counter()
    for i in range(steps):
        y = dn.denoise(imgs[i % 2], patch_size=1024, padding=500)
    torch.cuda.synchronize()
    e2e_ms = ctx.maxval((time.perf_counter() - t0) * 1e3)
    clocks = clk.finish()
    mpx = size * size / 1e6
    tf_img = 29.458 * (size / 4096) ** 2                # 29.46 TFLOP per 4096^2 image with 1024/500 patches
    return dict(workload=f'UDenoiseNet unet, Denoise.denoise({size}x{size}, patch_size=1024, padding=500), one image per step per GPU',
                metric='Mpx/s denoised', n_gpus=ctx.world, scaling='weak', steps=steps,
                value=ctx.world * steps * mpx / (dev_ms / 1e3), ms_per_image=dev_ms / steps,
                tflops_algorithmic=ctx.world * steps * tf_img / (dev_ms / 1e3),
                e2e=dict(value=ctx.world * steps * mpx / (e2e_ms / 1e3), ms_per_image=e2e_ms / steps, unit='Mpx/s',
                         h2d_bytes_per_step=size * size * 4, d2h_bytes_per_step=size * size * 4),
                kernel_launches_per_image_eager=launches, finite=bool(np.isfinite(y).all()), unit='Mpx/s', clocks=clocks)


def cfg4_train(ctx: Ctx, steps: int = 40, bn: bool = False, units: int = 32, weak: bool = False):
    """weak=True: 256 crops PER GPU (global minibatch 256 x world) -- the step's work per rank stays that of the single-GPU
    step, so the ratio to the N=1 time isolates what the collectives and the global loss cost."""
    import torch.nn as nn
    from common import gold, weights_of, seeded_state
    from topaz_b200 import ops
    from topaz_b200.methods import GE_binomial
    from topaz_b200.model.factory import get_feature_extractor
    from topaz_b200.model.classifier import LinearClassifier
    m = LinearClassifier(get_feature_extractor('resnet8', units=units, bn=bn))
    if bn:      # the default `topaz train` model (BatchNorm on, no packaged weights): seeded He init
        _load(m, seeded_state({k: tuple(v.shape) for k, v in m.state_dict().items()}, 401))
    else:
        _load(m, weights_of(gold('resnet8_u32_pretrained' if units == 32 else 'resnet8_u64_pretrained')))
    m.cuda(); m.train()
    tr = GE_binomial(m, torch.optim.Adam(m.parameters(), lr=2e-4), nn.BCEWithLogitsLoss(), 0.035)
    B = 256 * (ctx.world if weak else 1)
    b = B // ctx.world
    Y = torch.tensor([1.0] * (B // 16) + [0.0] * (B - B // 16), dtype=torch.float64)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0))       # spread positives over shards
    Xs = [torch.from_numpy(np.random.default_rng(4000 + s).standard_normal((B, 71, 71)).astype(np.float32))[perm][ctx.rank * b:(ctx.rank + 1) * b].cuda()
          for s in range(4)]
    Yl = Y[perm][ctx.rank * b:(ctx.rank + 1) * b].cuda()
    for s in range(6):
        out = tr.step(Xs[s % 4], Yl)
    ctx.sync()
    clk = ClockSampler(ctx.local); clk.start()
    l0 = ops.LAUNCH_COUNT
    t0 = time.perf_counter()
    for s in range(steps):
        out = tr.step(Xs[s % 4], Yl)
    torch.cuda.synchronize()
    ms = ctx.maxval((time.perf_counter() - t0) * 1e3)
    launches = (ops.LAUNCH_COUNT - l0) / steps
    clocks = clk.finish()
    coll = None
    if ctx.world > 1:      # the step's collectives in isolation (device-timed, max over ranks)
        from topaz_b200 import train_engine
        fp = train_engine.flat_params(m)
        g = torch.zeros_like(fp.flat_g)
        sc = torch.zeros(b * 3, dtype=torch.float32, device='cuda')
        gs = torch.zeros(ctx.world * b * 3, dtype=torch.float32, device='cuda')

        def timed(fn, n=20):
            for _ in range(3):
                fn()
            ctx.sync()
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            z.record(); torch.cuda.synchronize()
            return ctx.maxval(a.elapsed_time(z) / n * 1e3)
        coll = dict(grad_allreduce_us=timed(lambda: ctx.dist.all_reduce(g)), grad_bytes=int(g.numel() * 4),
                    logits_allgather_us=timed(lambda: ctx.dist.all_gather_into_tensor(gs, sc)), logits_bytes=int(gs.numel() * 4))
    return dict(workload=f'GE_binomial.step, resnet8_u{units}' + (' + BatchNorm (training mode)' if bn else '') +
                f', global minibatch {B} crops of 71x71 = {b} per GPU, Adam, incl. the per-step 5-float host read-back',
                metric='crops/s', unit='crops/s', n_gpus=ctx.world, scaling='weak' if weak else 'strong', steps=steps, value=steps * B / (ms / 1e3),
                ms_per_step=ms / steps, kernel_launches_per_step=launches, collectives=coll,
                tflops_algorithmic=steps * B * 3 * (59.27e-6 if units == 32 else 230.2e-6) / (ms / 1e3),
                last_out=[float(v) for v in out], clocks=clocks)


def cfg5_denoise3d(ctx: Ctx, size: int = 512):
    from common import gold, weights_of
    from topaz_b200 import engine
    from topaz_b200.denoising.models import UDenoiseNet3D
    from topaz_b200.denoise import Denoise3D
    from topaz_b200.parallel import shard_range
    d3 = Denoise3D(_load(UDenoiseNet3D(base_width=7), weights_of(gold('unet3d_pretrained_10a'))))
    S = size
    tomo = np.random.default_rng(5000).standard_normal((S, S, S)).astype(np.float32)
    npatch = int(np.ceil(S / 96)) ** 3
    lo, hi = shard_range(npatch, ctx.rank, ctx.world)
    d3.denoise(tomo[:200, :200, :200].copy(), verbose=False, patch_range=(0, 2))       # warm-up: two 192^3 patches
    ctx.sync()
    clk = ClockSampler(ctx.local); clk.start()
    t0 = time.perf_counter()
    y = d3.denoise(tomo, patch_size=96, padding=48, verbose=False, patch_range=(lo, hi))
    torch.cuda.synchronize()
    ms = ctx.maxval((time.perf_counter() - t0) * 1e3)
    clocks = clk.finish()
    return dict(workload=f'UDenoiseNet3D unet-3d-10a (pretrained), Denoise3D.denoise({S}^3, patch_size=96, padding=48): {npatch} patches of 192^3 '
                f'sharded over {ctx.world} GPU(s), host numpy in / out', metric='Mvox/s denoised', unit='Mvox/s', n_gpus=ctx.world,
                scaling='strong', value=S ** 3 / 1e6 / (ms / 1e3), ms_total=ms, patches_per_gpu=hi - lo,
                ms_per_patch=ms / max(1, hi - lo), tflops_algorithmic=npatch * 4.784 / (ms / 1e3), precision=engine.PRECISION,
                e2e=dict(value=S ** 3 / 1e6 / (ms / 1e3), unit='Mvox/s', h2d_bytes_per_step=S ** 3 * 4, d2h_bytes_per_step=S ** 3 * 4),
                finite=bool(np.isfinite(y).all()), clocks=clocks)


def gpu_library_baseline(ctx: Ctx, with_3d: bool = True):
    """The reference's networks as torch + cuDNN op sequences on this GPU (TF32 convolutions = torch's default and the
    reference's own GPU path; cudnn.benchmark off as in the reference, and on for fairness)."""
    import torch.nn.functional as F
    import torch_cudnn_baseline as T
    clk = ClockSampler(ctx.local); clk.start()
    out = dict(torch=torch.__version__, cudnn=torch.backends.cudnn.version(), allow_tf32=bool(torch.backends.cudnn.allow_tf32))
    torch.manual_seed(0)

    def best(make, x, key, per, reps=3):
        res = {}
        for flag in (False, True):
            torch.backends.cudnn.benchmark = flag
            try:
                with torch.no_grad():
                    f = make()
                    ms = T.time_it(lambda: f(x), warm=2, reps=reps)
                res['cudnn_benchmark_on' if flag else 'cudnn_benchmark_off'] = ms
            except Exception as e:      # e.g. out of memory in a cuDNN workspace
                res['error'] = f'{type(e).__name__}: {str(e)[:120]}'
            torch.cuda.empty_cache()
        torch.backends.cudnn.benchmark = False
        ok = [v for k, v in res.items() if k.startswith('cudnn')]
        if ok:
            res['ms'] = min(ok)
            res[per[0]] = per[1] / (min(ok) / 1e3)
        out[key] = res
    best(lambda: T.resnet8_dense(64), torch.randn(1, 1, 4096, 4096, device='cuda'), 'resnet8_u64_dense_4096', ('mpx_s', 16.777216))
    best(lambda: T.unet(48, 11, 5, 2), torch.randn(1, 1, 2048, 2048, device='cuda'), 'unet2d_2048_patch', ('mpx_s', 4.194304))
    if with_3d:
        best(lambda: T.unet(48, 7, 3, 3), torch.randn(1, 1, 192, 192, 192, device='cuda'), 'unet3d_192_patch', ('mvox_s', 7.077888), reps=2)
    X = torch.randn(256, 71, 71, device='cuda'); Y = torch.zeros(256, device='cuda'); Y[:16] = 1
    for units, bn in ((32, False), (32, True), (64, False)):
        net = T.TrainNet(units, bn=bn).cuda()
        opt = torch.optim.Adam(net.parameters(), lr=2e-4)

        def step():
            s = net(X)
            loss = F.binary_cross_entropy_with_logits(s[Y == 1], Y[Y == 1]) + torch.sigmoid(s[Y == 0]).sum() * 1e-3
            loss.backward()
            opt.step(); opt.zero_grad()
            return loss.item()      # the reference syncs every step (methods.py:148-165)
        ms = T.time_it(step, warm=5, reps=20)
        out[f'train_step_u{units}' + ('_bn' if bn else '')] = dict(ms=ms, crops_s=256 / (ms / 1e3), note='simplified loss (BCE + sigmoid sum): the GE term adds ~25 ATen launches and a CPU scipy call in the reference')
        del net, opt
    torch.cuda.empty_cache()
    out['clocks'] = clk.finish()
    return out
