#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs 3-5) on one GPU or sharded over ranks (torchrun):
  denoise   : UDenoiseNet `unet` via Denoise.denoise(x, patch_size=1024, padding=500) on 4096x4096 images
  train     : GE_binomial.step, resnet8_u32, 256 crops of 71x71 (batch sharded over ranks + NCCL all-reduce)
  denoise3d : UDenoiseNet3D via Denoise3D.denoise(patch 96, padding 48) on an S^3 tomogram (patches sharded)
One JSON line per workload on rank 0.  Weights: pretrained 2-D unet / resnet8_u32 from the golden fixtures;
the 3-D model uses seeded random weights (the 11.7 MB pretrained file is not shipped)."""
import argparse, json, os, sys, time
os.environ.setdefault('TPZ_X', '1'); os.environ['NCCL_DEBUG'] = 'WARN'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import torch.nn as nn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workloads', default='denoise,train,denoise3d')
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--tomo', type=int, default=288)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from common import gold, weights_of, seeded_state
    from common_shapes import unet_shapes
    from topaz_b200 import ops
    from topaz_b200.parallel import shard_range

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxms(ms):
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for wl in args.workloads.split(','):
        if wl == 'denoise':
            from topaz_b200.denoising.models import UDenoiseNet
            from topaz_b200.denoise import Denoise
            m = UDenoiseNet(base_width=11, top_width=5)
            m.load_state_dict({k: torch.from_numpy(v) for k, v in weights_of(gold('unet_pretrained')).items()})
            dn = Denoise(m)
            imgs = [(10 + 3 * np.random.default_rng(3000 + rank * 100 + i).standard_normal((4096, 4096))).astype(np.float32) for i in range(2)]
            dn.denoise(imgs[0], patch_size=1024, padding=500)
            sync(); l0 = ops.LAUNCH_COUNT; t0 = time.perf_counter()
            for i in range(args.steps):
                y = dn.denoise(imgs[i % 2], patch_size=1024, padding=500)
            torch.cuda.synchronize(); ms = maxms((time.perf_counter() - t0) * 1e3)
            xd = torch.from_numpy(imgs[0]).cuda()
            dn.denoise_patches_device(xd, 1024, 500); sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.steps):
                yd = dn.denoise_patches_device(xd, 1024, 500)
            e1.record(); torch.cuda.synchronize()
            dev_ms = maxms(e0.elapsed_time(e1))
            if rank == 0:
                print(json.dumps(dict(workload='UDenoiseNet unet denoise_patches_device(4096x4096, patch 1024, padding 500), device-resident',
                                      n_gpus=world, ms_per_image=dev_ms / args.steps, mpx_s=world * args.steps * 16.777216 / (dev_ms / 1e3),
                                      tflops=world * args.steps * 29.458 / (dev_ms / 1e3))))
                print(json.dumps(dict(workload='UDenoiseNet unet Denoise.denoise(4096x4096, patch 1024, padding 500), host numpy in/out',
                                      n_gpus=world, ms_per_image=ms / args.steps, mpx_s=world * args.steps * 16.777216 / (ms / 1e3),
                                      tflops=world * args.steps * 29.458 / (ms / 1e3), launches_per_image=(ops.LAUNCH_COUNT - l0) / args.steps,
                                      finite=bool(np.isfinite(y).all()))))
        elif wl in ('train', 'train_tf32', 'train_bn', 'train_u64'):
            from topaz_b200 import train_engine as _T
            _T.set_tf32(wl == 'train_tf32')
            from topaz_b200.methods import GE_binomial
            from topaz_b200.model.factory import get_feature_extractor
            from topaz_b200.model.classifier import LinearClassifier
            units = 64 if wl == 'train_u64' else 32
            m = LinearClassifier(get_feature_extractor('resnet8', units=units, bn=wl == 'train_bn'))
            if wl == 'train_bn':      # the default `topaz train` model (BatchNorm on, no packaged weights): seeded He init
                m.load_state_dict({k: torch.from_numpy(v) for k, v in seeded_state({k: tuple(v.shape) for k, v in m.state_dict().items()}, 401).items()})
            else:
                m.load_state_dict({k: torch.from_numpy(v) for k, v in weights_of(gold(f'resnet8_u{units}_pretrained')).items()})
            m.cuda(); m.train()
            tr = GE_binomial(m, torch.optim.Adam(m.parameters(), lr=2e-4), nn.BCEWithLogitsLoss(), 0.035)
            B = 256; b = B // world
            Y = torch.tensor([1.0] * 16 + [0.0] * 240, dtype=torch.float64)
            perm = torch.randperm(B, generator=torch.Generator().manual_seed(0))       # spread positives over shards
            Xs = [torch.from_numpy(np.random.default_rng(4000 + s).standard_normal((B, 71, 71)).astype(np.float32))[perm][rank * b:(rank + 1) * b].cuda() for s in range(4)]
            Yl = Y[perm][rank * b:(rank + 1) * b].cuda()
            for s in range(5):
                out = tr.step(Xs[s % 4], Yl)
            sync(); l0 = ops.LAUNCH_COUNT; t0 = time.perf_counter()
            n = 50
            for s in range(n):
                out = tr.step(Xs[s % 4], Yl)
            torch.cuda.synchronize(); ms = maxms((time.perf_counter() - t0) * 1e3)
            if rank == 0:
                print(json.dumps(dict(workload='GE_binomial.step resnet8_u32, global minibatch 256 crops 71x71 (incl. per-step host readback)' + (', single-pass TF32 mode' if wl == 'train_tf32' else ', 3xTF32') + (', BatchNorm (training mode)' if wl == 'train_bn' else ''),
                                      n_gpus=world, ms_per_step=ms / n, crops_s=n * B / (ms / 1e3), launches_per_step=(ops.LAUNCH_COUNT - l0) / n,
                                      last_out=out)))
            _T.set_tf32(False)
        elif wl == 'train_e2e':
            # GE-binomial training fed by the GPU crop sampler (the reference's host loader: 0.28-0.35 s per minibatch)
            from topaz_b200.methods import GE_binomial
            from topaz_b200.sampler import GpuCropSampler
            from topaz_b200.model.factory import get_feature_extractor
            from topaz_b200.model.classifier import LinearClassifier
            m = LinearClassifier(get_feature_extractor('resnet8', units=32, bn=False))
            m.load_state_dict({k: torch.from_numpy(v) for k, v in weights_of(gold('resnet8_u32_pretrained')).items()})
            m.cuda(); m.train()
            tr = GE_binomial(m, torch.optim.Adam(m.parameters(), lr=2e-4), nn.BCEWithLogitsLoss(), 0.035)
            g = np.random.default_rng(7 + rank)
            mics = [g.standard_normal((1024, 1024)).astype(np.float32) for _ in range(16)]
            pos = []
            for k in range(16):
                for (py, px) in g.integers(40, 984, size=(60, 2)):
                    for dy in range(-3, 4):
                        for dx in range(-3, 4):
                            if dy * dy + dx * dx <= 9:
                                pos.append((k, py + dy, px + dx))
            smp = GpuCropSampler([mics], np.array(pos, dtype=np.int32), 71, positive_balance=0.0625, seed=rank)
            b = 256 // world
            for _ in range(5):
                tr.step(*smp.sample(b))
            sync(); t0 = time.perf_counter(); n = 50
            for _ in range(n):
                out = tr.step(*smp.sample(b))
            torch.cuda.synchronize(); ms = maxms((time.perf_counter() - t0) * 1e3)
            if rank == 0:
                print(json.dumps(dict(workload='GPU crop sampler (rotate+flip, 16 micrographs 1024^2) + GE_binomial.step, 256 crops/minibatch',
                                      n_gpus=world, ms_per_step=ms / n, crops_s=n * 256 / (ms / 1e3), last_out=out)))
        elif wl == 'denoise3d':
            from topaz_b200.denoising.models import UDenoiseNet3D
            from topaz_b200.denoise import Denoise3D
            m = UDenoiseNet3D(nf=48, base_width=7, top_width=3)
            m.load_state_dict({k: torch.from_numpy(v) for k, v in seeded_state(unet_shapes(48, 7, 3, 3), 202).items()})
            d3 = Denoise3D(m)
            S = args.tomo
            tomo = np.random.default_rng(5000).standard_normal((S, S, S)).astype(np.float32)
            npatch = int(np.ceil(S / 96)) ** 3
            lo, hi = shard_range(npatch, rank, world)
            d3.denoise(tomo[:96, :96, :96].copy(), verbose=False)       # warm-up: one 192^3 patch
            sync(); t0 = time.perf_counter()
            y = d3.denoise(tomo, patch_size=96, padding=48, verbose=False, patch_range=(lo, hi))
            torch.cuda.synchronize(); ms = maxms((time.perf_counter() - t0) * 1e3)
            if rank == 0:
                print(json.dumps(dict(workload=f'UDenoiseNet3D Denoise3D.denoise({S}^3, patch 96, padding 48), {npatch} patches of 192^3, host numpy in/out',
                                      n_gpus=world, ms_total=ms, ms_per_patch=ms / max(1, hi - lo), mvox_s=S ** 3 / 1e6 / (ms / 1e3),
                                      tflops=npatch * 4.784 / (ms / 1e3), finite=bool(np.isfinite(y).all()))))
        elif wl == 'preprocess':
            # `topaz preprocess -s 8`: Fourier-crop downsample + GMM normalisation of a K3-sized micrograph; then NMS timing
            from topaz_b200 import preprocess, stats
            from topaz_b200.algorithms import non_maximum_suppression
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            from oracle import topaz_oracle as O
            gq = np.random.default_rng(6000 + rank)
            big = (100 + 5 * gq.standard_normal((7676, 7420)) + 10 * (gq.random((7676, 7420)) < 0.15)).astype(np.float32)
            xd = torch.from_numpy(big).cuda()
            preprocess.downsample_device(xd, 8); torch.cuda.synchronize()      # builds + caches the operators
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                sm = preprocess.downsample_device(xd, 8)
            e1.record(); torch.cuda.synchronize(); ds_ms = e0.elapsed_time(e1) / 5
            t0 = time.perf_counter(); small_h = preprocess.downsample(big, 8); ds_host_ms = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter(); ref = O.downsample(big, 8); cpu_ds_ms = (time.perf_counter() - t0) * 1e3
            err = float(np.abs(small_h - ref).max() / np.abs(ref).max())
            stats.normalize_device(sm, method='gmm'); torch.cuda.synchronize()
            t0 = time.perf_counter(); yn, md = stats.normalize_device(sm, method='gmm'); torch.cuda.synchronize()
            gmm_ms = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter(); yr, mu, std, pi = O.gmm_normalize(sm.cpu().numpy()); cpu_gmm_ms = (time.perf_counter() - t0) * 1e3
            full = torch.from_numpy(big[:4096, :4096].copy()).cuda()
            stats.normalize_device(full, method='gmm', alpha=2, beta=2); torch.cuda.synchronize()
            t0 = time.perf_counter(); _, md2 = stats.normalize_device(full, method='gmm', alpha=2, beta=2); torch.cuda.synchronize()
            gmm_full_ms = (time.perf_counter() - t0) * 1e3
            if rank == 0:
                print(json.dumps(dict(workload='preprocess: downsample(7676x7420, 8) -> 959x927, then stats.normalize(method=gmm)',
                                      downsample_device_ms=ds_ms, downsample_host_in_out_ms=ds_host_ms, cpu_oracle_downsample_ms=cpu_ds_ms,
                                      downsample_max_rel_err=err, gmm_normalize_ms=gmm_ms, cpu_oracle_gmm_ms=cpu_gmm_ms,
                                      mu=[md['mu'], mu], std=[md['std'], std], gmm_normalize_4096sq_ms=gmm_full_ms, pi_4096=md2['pi'])))
        elif wl == 'nms':
            from topaz_b200.algorithms import non_maximum_suppression, non_maximum_suppression_3d
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            from oracle import topaz_oracle as O
            gq = np.random.default_rng(7000)
            sc = gq.standard_normal((4096, 4096)).astype(np.float32)
            sd = torch.from_numpy(sc).cuda()
            res = {}
            for r, thr in [(8, 1.5), (8, -6.0), (14, 0.0)]:
                non_maximum_suppression(sd, r, thr); torch.cuda.synchronize()
                t0 = time.perf_counter(); s, c = non_maximum_suppression(sd, r, thr); res[f'2d_r{r}_thr{thr}_ms'] = (time.perf_counter() - t0) * 1e3
                res[f'2d_r{r}_thr{thr}_picks'] = len(s)
            t0 = time.perf_counter(); s_ref, c_ref = O.nms(sc, 8, 1.5); res['cpu_oracle_2d_r8_thr1.5_ms'] = (time.perf_counter() - t0) * 1e3
            vol = torch.from_numpy(gq.standard_normal((128, 512, 512)).astype(np.float32)).cuda()
            non_maximum_suppression_3d(vol, 6, threshold=1.5); torch.cuda.synchronize()
            t0 = time.perf_counter(); s3, c3 = non_maximum_suppression_3d(vol, 6, threshold=1.5); res['3d_128x512x512_r6_thr1.5_ms'] = (time.perf_counter() - t0) * 1e3
            res['3d_picks'] = len(s3)
            if rank == 0:
                print(json.dumps(dict(workload='greedy NMS on 4096^2 score maps / a 128x512x512 volume (host-visible time incl. result read-back)', **res)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
