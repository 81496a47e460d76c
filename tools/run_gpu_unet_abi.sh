# First hardware run of the U-Net entry points of the model-level C ABI (written after round 2's GPU budget was spent).
#   1. the bit-for-bit tests against the Python plans + the plain-C host (tests/test_gpu_unet_abi.py; --runxfail turns the
#      non-strict xfail marks into plain pass / fail)
#   2. A/B of the host cost the handle removes: one 2048^2 patch, eager (no CUDA graph), Python plans vs C handle
timeout 900 python -m pytest tests/test_gpu_unet_abi.py -m gpu -q -p no:cacheprovider --runxfail 2>&1 | tail -5 | cut -c1-220
for eng in py c; do
TPZ_UNET_ENGINE=$eng TPZ_DENOISE_GRAPH=0 timeout 300 python - <<'PY'
import os, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from common import gold, weights_of
from topaz_b200.denoising.models import UDenoiseNet
from topaz_b200 import engine, ops
g = gold('unet_pretrained')
m = UDenoiseNet(base_width=11, top_width=5)
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in weights_of(g).items()}); m.eval().cuda()
x = torch.randn(1, 1, 2048, 2048, device='cuda')
with torch.no_grad():
    for _ in range(3):
        y = engine.unet_forward(m, x)
    torch.cuda.synchronize()
    l0 = ops.LAUNCH_COUNT
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = engine.unet_forward(m, x)
    e1.record()
    host_ms = (time.perf_counter() - t0) * 100          # host time to ENQUEUE one forward (before the sync)
    torch.cuda.synchronize()
print(f"TPZ_UNET_ENGINE={os.environ['TPZ_UNET_ENGINE']}: {e0.elapsed_time(e1) / 10:.2f} ms / patch on the device, {host_ms:.2f} ms host enqueue, "
      f"{(ops.LAUNCH_COUNT - l0) // 10} launches, checksum {float(y.double().sum()):.6f}")
PY
done
