# Round 2, GPU call Z: 256-row stage of the 64-channel halo kernel; resnet8_u64 training step (cfg4 secondary) vs torch+cuDNN in the same line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3 | cut -c1-300
timeout 600 python bench.py --steps 3 --extras cfg4,cfg4bn,cfg4u64,lib --no-cpu-baseline 2>gpurun_out/r2z_bench.err > gpurun_out/r2z_bench.json; python - <<PY
import json
d=json.load(open("gpurun_out/r2z_bench.json"))
for k,v in d["extra"].items():
    if k.startswith("cfg4"): print(k, v.get("ms_per_step"), v.get("kernel_launches_per_step"))
lib=d["extra"].get("gpu_library_baseline",{})
print({k:(v.get("ms") if isinstance(v,dict) else v) for k,v in lib.items() if k.startswith("train")})
print(d["extra"].get("speedup_vs_gpu_library"))
PY
