#!/bin/bash
# usage: bash tools/run_scaling.sh N  (under gpurun --gpus N): headline bench + DP training check at N GPUs
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 300 $TR tools/dp_check.py 2>&1 | grep dp_check
timeout 600 $TR bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_n$N.err > gpurun_out/bench_n$N.json
python - <<PY
import json
for l in open("gpurun_out/bench_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l); print("bench N=$N", round(d["value"],1), "Mpx/s  ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), d["clocks"])
PY
timeout 600 $TR tools/bench_extra.py --workloads train,denoise,denoise3d --tomo 384 2>gpurun_out/extra_n$N.err | tee gpurun_out/bench_extra_n$N.jsonl | cut -c1-300
