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
import sys
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


sys.path.insert(0, os.path.join(ROOT, 'tools'))
from workloads import ClockSampler, Ctx, cfg3_denoise2d, cfg4_train, cfg5_denoise3d, gpu_library_baseline  # noqa: E402


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
        dt = None
        for _ in range(2):            # best of two timed calls: a single call let scheduler noise pick the thread count
            t0 = time.perf_counter(); fn(); d = time.perf_counter() - t0
            dt = d if dt is None or d < dt else dt
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
    ap.add_argument('--extras', default='cfg3,cfg4,cfg5,lib',
                    help='secondary BASELINE workloads folded into the line\'s `extra` block (comma list of cfg3,cfg4,cfg4bn,cfg4u64,cfg5,lib; "none" to skip)')
    ap.add_argument('--tomo', type=int, default=512, help='cfg5 tomogram edge (512 = the full BASELINE size: 216 patches of 192^3)')
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
    # default engine = the model-level C ABI (tpz_model_*): it records the event pair around the dominant launch itself
    c_model = (model.__dict__.get('_tpz_plans', {}).get('dense_c') or (None, None, None))[2]
    if c_model is not None:
        c_model.timing(True)
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
    if c_model is not None:
        dom_ms += c_model.timing_read()
        c_model.timing(False)
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

    # --- secondary workloads (BASELINE configs 3-5) and the same-GPU library baseline, each with its own clock sample ---
    extra = {}
    wanted = [] if args.extras == 'none' else [w for w in args.extras.split(',') if w]
    ctx = Ctx()
    del devimgs, host
    torch.cuda.empty_cache()
    for name in wanted:
        try:
            if name == 'cfg3':
                extra['cfg3_unet2d_denoise'] = cfg3_denoise2d(ctx, steps=4)
            elif name == 'cfg4':
                extra['cfg4_ge_binomial_train'] = cfg4_train(ctx, steps=40)
                if world > 1:      # the same step at 256 crops per GPU: separates the collectives' cost from the shrinking shard
                    try:
                        extra['cfg4_ge_binomial_train_weak'] = cfg4_train(ctx, steps=40, weak=True)
                    except Exception as e:
                        extra['cfg4_ge_binomial_train_weak_error'] = f'{type(e).__name__}: {str(e)[:300]}'
            elif name == 'cfg4bn':
                extra['cfg4_ge_binomial_train_bn'] = cfg4_train(ctx, steps=40, bn=True)
            elif name == 'cfg4u64':
                extra['cfg4_ge_binomial_train_u64'] = cfg4_train(ctx, steps=20, units=64)
            elif name == 'cfg5':
                extra['cfg5_unet3d_denoise'] = cfg5_denoise3d(ctx, size=args.tomo)
            elif name == 'lib' and world == 1:
                extra['gpu_library_baseline'] = gpu_library_baseline(ctx)
        except Exception as e:      # a failing secondary workload must not lose the headline line
            extra[name + '_error'] = f'{type(e).__name__}: {str(e)[:300]}'
        torch.cuda.empty_cache()
    lib = extra.get('gpu_library_baseline')

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
        if lib:      # the north-star target: >= 1.5x the reference's own torch/cuDNN path on the same B200, same run
            sp = {}
            if 'ms' in lib.get('resnet8_u64_dense_4096', {}):
                sp['resnet8_u64_scoring_4096'] = lib['resnet8_u64_dense_4096']['ms'] / (ms_max / args.steps)
            c3 = extra.get('cfg3_unet2d_denoise')
            if c3 and 'ms' in lib.get('unet2d_2048_patch', {}):
                # Denoise.denoise(4096^2, 1024/500) = 4 crops of 2048^2 + 8 of 2048x1524 + 4 of 1524^2 = 51.04 Mpx processed
                sp['unet2d_denoise_4096'] = (lib['unet2d_2048_patch']['ms'] * 51.04 / 4.194304) / c3['ms_per_image']
            c4 = extra.get('cfg4_ge_binomial_train')
            if c4 and 'train_step_u32' in lib:
                sp['ge_binomial_train_step'] = lib['train_step_u32']['ms'] / c4['ms_per_step']
            c4b = extra.get('cfg4_ge_binomial_train_bn')
            if c4b and 'train_step_u32_bn' in lib:
                sp['ge_binomial_train_step_bn'] = lib['train_step_u32_bn']['ms'] / c4b['ms_per_step']
            c4u = extra.get('cfg4_ge_binomial_train_u64')
            if c4u and 'train_step_u64' in lib:
                sp['ge_binomial_train_step_u64'] = lib['train_step_u64']['ms'] / c4u['ms_per_step']
            c5 = extra.get('cfg5_unet3d_denoise')
            if c5 and 'ms' in lib.get('unet3d_192_patch', {}):
                sp['unet3d_denoise_patch'] = lib['unet3d_192_patch']['ms'] / c5['ms_per_patch']
            extra['speedup_vs_gpu_library'] = sp
        if extra:
            line['extra'] = extra
        if not args.no_cpu_baseline and world == 1:      # reported on rank 0 at N=1 only (the reference arm covers N>1)
            v, dt, thr = cpu_oracle_mpxs()
            line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': thr, 'kind': 'port',
                                    'sample': f'one 512x512 micrograph x3 (bounded sample), oracle torch CPU fp32, {dt:.2f} s each, {thr} threads = fastest of the tried counts on {os.cpu_count()} logical CPUs'}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
