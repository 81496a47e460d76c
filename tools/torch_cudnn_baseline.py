#!/usr/bin/env python
"""Same-GPU library baseline: the reference's networks are plain torch.nn stacks that dispatch to cuDNN/ATen
(TF32 convs enabled by default).  The reference itself cannot travel to the GPU box, so this tool rebuilds
the same op sequences with torch.nn.functional (no oracle import, no topaz_b200 kernels) and times them with
CUDA events: (1) ResNet8-u64 dense forward on one 4096x4096 micrograph, (2) UDenoiseNet forward on one
2048x2048 patch, (3) a GE-binomial-style training step (fwd + bwd + Adam) on 256 crops.  Random weights —
only the timing matters.  Output: one JSON line per workload (informational)."""
import json
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F


def time_it(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


def resnet8_dense(u=64):
    dev = 'cuda'
    mk = lambda co, ci, k: (torch.randn(co, ci, k, k, device=dev) * (2.0 / (ci * k * k)) ** 0.5, torch.zeros(co, device=dev))
    W = dict(c7=mk(u, 1, 7), r1a=mk(u, u, 3), r1b=mk(u, u, 3), r2a=mk(u, u, 3), r2b=mk(2 * u, u, 3), r2p=mk(2 * u, u, 1),
             r3a=mk(2 * u, 2 * u, 3), r3b=mk(2 * u, 2 * u, 3), c5=mk(4 * u, 2 * u, 5), cls=mk(1, 4 * u, 1))

    def fwd(x):
        x = F.pad(x, (35,) * 4)
        x = F.relu(F.conv2d(x, *W['c7']))
        def resid(x, a, b, d0, d1, proj=None):
            h = F.relu(F.conv2d(x, *a, dilation=d0))
            y = F.conv2d(h, *b, dilation=d1)
            e = d0 + d1
            s = x[:, :, e:-e, e:-e]
            if proj is not None:
                s = F.conv2d(s, proj[0])
            return F.relu(y + s)
        x = resid(x, W['r1a'], W['r1b'], 2, 4)
        x = resid(x, W['r2a'], W['r2b'], 2, 4, W['r2p'])
        x = resid(x, W['r3a'], W['r3b'], 4, 8)
        x = F.relu(F.conv2d(x, *W['c5'], dilation=4))
        return F.conv2d(x, *W['cls'])
    return fwd


def unet(nf=48, base=11, top=5, dims=2):
    dev = 'cuda'
    Conv = nn.Conv3d if dims == 3 else nn.Conv2d
    pool = F.max_pool3d if dims == 3 else F.max_pool2d
    def cv(ci, co, k):
        c = Conv(ci, co, k, padding=k // 2).to(dev); return c
    enc = [cv(1, nf, base)] + [cv(nf, nf, 3) for _ in range(5)]
    dec = {5: (cv(2 * nf, 2 * nf, 3), cv(2 * nf, 2 * nf, 3))}
    for l in (4, 3, 2):
        dec[l] = (cv(3 * nf, 2 * nf, 3), cv(2 * nf, 2 * nf, 3))
    d1 = (cv(2 * nf + 1, 64, top), cv(64, 32, top), cv(32, 1, top))

    def fwd(x):
        skips = [x]; h = x
        for i, c in enumerate(enc):
            h = F.leaky_relu(c(h), 0.1)
            if i < 5:
                h = pool(h, 2); skips.append(h)
        for l in (5, 4, 3, 2):
            s = skips[l - 1]
            h = torch.cat([F.interpolate(h, size=s.shape[2:], mode='nearest'), s], 1)
            h = F.leaky_relu(dec[l][1](F.leaky_relu(dec[l][0](h), 0.1)), 0.1)
        h = torch.cat([F.interpolate(h, size=x.shape[2:], mode='nearest'), x], 1)
        return d1[2](F.leaky_relu(d1[1](F.leaky_relu(d1[0](h), 0.1)), 0.1))
    return fwd


class TrainNet(nn.Module):
    """ResNet8 in its strided training geometry; bn=True: BatchNorm after every conv / residual sum (the model `topaz train
    --no-pretrained` builds), convs without bias."""
    def __init__(s, u=32, bn=False):
        super().__init__()
        b = not bn
        s.c7 = nn.Conv2d(1, u, 7, stride=2, bias=b)
        s.r1a, s.r1b = nn.Conv2d(u, u, 3, bias=b), nn.Conv2d(u, u, 3, dilation=2, bias=b)
        s.r2a, s.r2b, s.r2p = nn.Conv2d(u, u, 3, bias=b), nn.Conv2d(u, 2 * u, 3, dilation=2, stride=2, bias=b), nn.Conv2d(u, 2 * u, 1, stride=2, bias=False)
        s.r3a, s.r3b = nn.Conv2d(2 * u, 2 * u, 3, bias=b), nn.Conv2d(2 * u, 2 * u, 3, dilation=2, bias=b)
        s.c5, s.cls = nn.Conv2d(2 * u, 4 * u, 5, bias=b), nn.Conv2d(4 * u, 1, 1)
        norm = (lambda c: nn.BatchNorm2d(c)) if bn else (lambda c: nn.Identity())
        s.n7, s.n1a, s.n1b, s.n2a, s.n2b, s.n3a, s.n3b, s.n5 = (norm(c) for c in (u, u, u, u, 2 * u, 2 * u, 2 * u, 4 * u))

    def forward(s, x):
        x = F.relu(s.n7(s.c7(x.unsqueeze(1))))
        x = F.relu(s.n1b(s.r1b(F.relu(s.n1a(s.r1a(x)))) + x[:, :, 3:-3, 3:-3]))
        x = F.relu(s.n2b(s.r2b(F.relu(s.n2a(s.r2a(x)))) + s.r2p(x[:, :, 3:-3, 3:-3])))
        x = F.relu(s.n3b(s.r3b(F.relu(s.n3a(s.r3a(x)))) + x[:, :, 3:-3, 3:-3]))
        return s.cls(F.relu(s.n5(s.c5(x)))).view(-1)


def main():
    torch.manual_seed(0)
    info = dict(torch=torch.__version__, cudnn=torch.backends.cudnn.version(), allow_tf32=torch.backends.cudnn.allow_tf32,
                gpu=torch.cuda.get_device_name(0))
    out = []
    with torch.no_grad():
        for bench_flag in (False, True):
            torch.backends.cudnn.benchmark = bench_flag
            f = resnet8_dense(64)
            x = torch.randn(1, 1, 4096, 4096, device='cuda')
            try:
                ms = time_it(lambda: f(x))
                out.append(dict(workload='resnet8_u64 dense 4096x4096 (torch+cuDNN, TF32 default)', cudnn_benchmark=bench_flag,
                                ms=ms, mpx_s=16.777216 / (ms / 1e3)))
            except Exception as e:
                out.append(dict(workload='resnet8_u64 dense 4096x4096', cudnn_benchmark=bench_flag, error=str(e)[:200]))
            torch.cuda.empty_cache()
            f = unet()
            x = torch.randn(1, 1, 2048, 2048, device='cuda')
            try:
                ms = time_it(lambda: f(x))
                out.append(dict(workload='UDenoiseNet 2048x2048 patch (torch+cuDNN)', cudnn_benchmark=bench_flag, ms=ms,
                                mpx_s=4.194304 / (ms / 1e3)))
            except Exception as e:
                out.append(dict(workload='UDenoiseNet 2048x2048', cudnn_benchmark=bench_flag, error=str(e)[:200]))
            torch.cuda.empty_cache()
    torch.backends.cudnn.benchmark = False
    X = torch.randn(256, 71, 71, device='cuda'); Y = torch.zeros(256, device='cuda'); Y[:16] = 1
    for bn in (False, True):
        net = TrainNet(32, bn=bn).cuda()
        opt = torch.optim.Adam(net.parameters(), lr=2e-4)

        def step():
            s = net(X)
            loss = F.binary_cross_entropy_with_logits(s[Y == 1], Y[Y == 1]) + torch.sigmoid(s[Y == 0]).sum() * 1e-3
            loss.backward()
            opt.step(); opt.zero_grad()
            return loss.item()      # the reference syncs every step (methods.py:148-165)
        ms = time_it(step, warm=5, reps=20)
        out.append(dict(workload='GE-style train step, resnet8_u32' + (' + BatchNorm' if bn else '') + ', 256x71x71 (torch+cuDNN, simplified loss)',
                        ms=ms, crops_s=256 / (ms / 1e3)))
    for o in out:
        o.update(info)
        print(json.dumps(o))


if __name__ == '__main__':
    main()
