#!/usr/bin/env python
"""Benchmark of the hot path (driver contract: one JSON line on rank 0).

Workload (BASELINE.json configs[1]): per-pixel scoring of synthetic 4096x4096 micrographs with the pretrained
ResNet-8 (units 64) classifier, dense ("filled") forward.  One step = one micrograph per GPU.
  value : megapixels/s with inputs resident in HBM (whole job, all GPUs), device-timed (CUDA events, max over ranks)
  e2e   : same metric through the public API with HOST buffers (pinned H2D + network + D2H per step)
  roofline : the dominant kernel (last 5x5 conv 128->256 fused with the 1x1 classifier), timed with CUDA events
             inside the timed region, algorithmic FLOPs / time vs the measured bf16/fp16 tensor peak
  cpu_baseline : the CPU oracle (torch fp32, all host threads) on a bounded sample (512x512)
`--impl reference` times the reference algorithm's CPU implementation (oracle port) instead.
Multi-GPU: images are sharded one-per-rank (weak scaling), no data-path collective.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

# The driver expects exactly ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner on
# fd 1 at NCCL_DEBUG=VERSION and above), so fd 1 is pointed at stderr for the whole run and the JSON line is written to
# the saved original descriptor at the end (emit()).
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)
sys.stdout = os.fdopen(os.dup(2), 'w', buffering=1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + '\n').encode())

import numpy as np
import torch

METRIC = 'megapixels/s scored (ResNet8-u64 dense forward, 4096x4096 micrographs)'
UNIT = 'Mpx/s'
# algorithmic FLOPs (2*MAC) of the reference network per OUTPUT pixel at 4096^2 (SURVEY 8d / BASELINE.md section 2)
FLOP_PER_PX_4096 = 2636177.0


def load_traffic():
    """DRAM bytes (read+write) per launch of the dominant kernel from the committed ncu --set full capture."""
    path = os.path.join(ROOT, 'profiles', 'dominant_kernel_ncu.json')
    if os.path.exists(path):
        return json.load(open(path)).get('dram_bytes_per_launch')
    return None


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops=float(p.get('bf16_tflops_sustained', p.get('bf16_tflops'))), hbm=float(p['hbm_gbs']),
                    src='measured (MEASURED_PEAKS.json, sustained bf16/fp16 dense)')
    return dict(tflops=1400.0, hbm=6650.0, src='fallback (B200_PROFILING.md)')


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons during the timed region.  Uses NVML in-process (initialised before the timed
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

    def summary(self):
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unavailable'])
        sm = sorted(r[0] for r in self.rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[2][i] for r in self.rows)]
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=self.rows[0][1], reasons=reasons, samples=len(self.rows),
                    source='nvml' if self.nvml is not None else 'nvidia-smi')


def pretrained_u64_state():
    from common import gold, weights_of
    return weights_of(gold('resnet8_u64_pretrained'))


def synth_image(i, size):
    return np.random.default_rng(1000 + i).standard_normal((size, size)).astype(np.float32)


def best_cpu_threads(fn):
    """oneDNN convolutions on a 128-thread host are often faster with fewer threads; give the CPU reference its best
    configuration: time one call at each candidate thread count and keep the fastest."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (ncpu, ncpu // 2, 32, 16, 8) if 1 <= c <= ncpu}, reverse=True)
    best, best_t = ncpu, None
    for c in cands:
        torch.set_num_threads(c)
        fn()
        t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_oracle_mpxs(size=512, reps=3, threads=None):
    from oracle import topaz_oracle as O
    sd = pretrained_u64_state()
    x = synth_image(0, size)[None, None]
    threads = threads or best_cpu_threads(lambda: O.classifier_forward(sd, x, 'resnet8', 64, filled=True))
    torch.set_num_threads(threads)
    O.classifier_forward(sd, x, 'resnet8', 64, filled=True)      # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        O.classifier_forward(sd, x, 'resnet8', 64, filled=True)
    dt = (time.perf_counter() - t0) / reps
    return size * size / 1e6 / dt, dt, threads


def run_reference(args, rank):
    """--impl reference: the reference algorithm on the host CPU (oracle port of the reference's own PyTorch
    CPU path; the reference itself cannot travel to the GPU box).  Each step = one bounded 512x512 sample."""
    if rank != 0:
        return
    size = 512
    from oracle import topaz_oracle as O
    sd = pretrained_u64_state()
    x = synth_image(0, size)[None, None]
    threads = best_cpu_threads(lambda: O.classifier_forward(sd, x, 'resnet8', 64, filled=True))
    for _ in range(max(1, min(args.warmup, 2))):
        O.classifier_forward(sd, x, 'resnet8', 64, filled=True)
    steps = max(1, min(args.steps, 8))
    t0 = time.perf_counter()
    for _ in range(steps):
        O.classifier_forward(sd, x, 'resnet8', 64, filled=True)
    dt = time.perf_counter() - t0
    v = steps * size * size / 1e6 / dt
    sample = f'{steps} x one {size}x{size} micrograph (bounded sample of the 4096x4096 workload), torch CPU fp32, {threads} threads (fastest of the tried thread counts; host has {os.cpu_count()} logical CPUs)'
    emit({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * dt / steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'resnet8_u64 dense scoring, CPU reference path', 'sample': sample},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--size', type=int, default=4096)
    ap.add_argument('--images', type=int, default=8, help='distinct synthetic micrographs cycled per rank (64 in the full job)')
    ap.add_argument('--variant', default='auto', choices=['auto', 'v1', 'v2'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    from topaz_b200 import ops
    from topaz_b200.model.classifier import LinearClassifier
    from topaz_b200.model.features.resnet import ResNet8
    from topaz_b200.extract import score_arrays
    ops.TC_VARIANT = args.variant

    model = LinearClassifier(ResNet8(units=64, bn=False))
    model.load_state_dict({k: torch.from_numpy(v) for k, v in pretrained_u64_state().items()})
    model.eval(); model.fill(); model.cuda()

    S = args.size
    nimg = max(1, args.images)
    host = [torch.from_numpy(synth_image(rank * 1000 + i, S)).pin_memory() for i in range(nimg)]
    devimgs = [h.cuda() for h in host]

    # --- event recorder for the dominant kernel (roofline) ---
    recorded = []

    def hook(tag):
        if tag != 'dominant':
            return None
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        recorded.append(ev)
        return ev

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(i):
        with torch.no_grad():
            return model(devimgs[i % nimg][None, None])

    for i in range(args.warmup):
        step_device(i)
    barrier()
    sampler = ClockSampler(local); sampler.start()
    ops.EVENT_HOOK = hook
    launches0 = ops.LAUNCH_COUNT
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        y = step_device(i)
    t1.record()
    barrier()
    ops.EVENT_HOOK = None
    launches = ops.LAUNCH_COUNT - launches0
    ms = t0.elapsed_time(t1)
    dom_ms = [a.elapsed_time(b) for a, b in recorded]
    checksum = float(y.double().sum().item())

    # --- end-to-end leg: host numpy in -> host numpy out through the public array API ---
    # warm-up long enough for the pinned-host caching allocator to hold every staging block the steady state needs
    # (a cudaHostAlloc of 64 MB inside the timed region costs tens of ms)
    for _ in score_arrays(model, [host[i % nimg].numpy() for i in range(max(4, args.warmup))], device=local):
        pass
    barrier()
    e0 = time.perf_counter()
    tcuda0, tcuda1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tcuda0.record()
    n_e2e = 0
    for out in score_arrays(model, (host[i % nimg].numpy() for i in range(args.steps)), device=local):
        n_e2e += 1
        assert out.shape == (S, S), f'end-to-end leg produced {out.shape}, expected {(S, S)} (same network as the timed leg)'
    tcuda1.record()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - e0
    sampler.stop_flag = True
    sampler.join(timeout=2)

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = t.tolist()
    mpx_step = S * S / 1e6
    value = world * args.steps * mpx_step / (ms_max / 1e3)
    e2e_value = world * n_e2e * mpx_step / (e2e_ms_max / 1e3)

    if rank == 0:
        peaks = load_peaks()
        # dominant kernel: last BasicConv 5x5 dil4 128->256 (+ fused 1x1 classifier): 2*Ho*Wo*Co*K flops per launch
        dom_flops = 2.0 * S * S * 256 * (128 * 25) + 2.0 * S * S * 256
        dom_avg_ms = sum(dom_ms) / max(1, len(dom_ms))
        achieved = dom_flops / (dom_avg_ms * 1e-3) / 1e12 if dom_ms else None
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f16 operands, f32 accumulate (same 11-bit significand as the reference GPU path\'s TF32)',
            'data': f'synthetic N(0,1) micrographs {S}x{S} (default_rng(1000+i)); pretrained resnet8_u64 weights',
            'config': {'workload': f'resnet8_u64 dense scoring of {S}x{S} micrographs, {nimg} distinct images cycled per GPU, one image per step per GPU',
                       'l2': 'per-layer activations are 2.2-4.4 GB >> 126 MB L2 (no flush needed)', 'tc_variant': args.variant,
                       'flop_per_step_T': FLOP_PER_PX_4096 * S * S / 1e12 if S == 4096 else None, 'checksum': checksum},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': S * S * 4, 'd2h_bytes_per_step': S * S * 4,
                    'ms_per_step': e2e_ms_max / max(1, n_e2e)},
            'gpu_launches': launches,
            'clocks': sampler.summary(),
            'roofline': {'bound': 'tensor', 'kernel': 'tc_conv (5x5 dil4 128->256 + fused classifier dot)',
                         'achieved': achieved, 'peak': peaks['tflops'], 'unit': 'TFLOP/s',
                         'frac': (achieved / peaks['tflops']) if achieved else None, 'traffic': load_traffic() if S == 4096 else None,
                         'peak_source': peaks['src'], 'ms_per_launch': dom_avg_ms, 'launches_timed': len(dom_ms),
                         'step_tflops': (FLOP_PER_PX_4096 * S * S / 1e12) / (ms_max / args.steps / 1e3) if S == 4096 else None},
        }
        if not args.no_cpu_baseline and world == 1:      # reported on rank 0 at N=1 only (the reference arm covers N>1)
            v, dt, thr = cpu_oracle_mpxs()
            line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': thr, 'kind': 'port',
                                    'sample': f'one 512x512 micrograph x3 (bounded sample), oracle torch CPU fp32, {dt:.2f} s each, {thr} threads = fastest of the tried counts on {os.cpu_count()} logical CPUs'}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
