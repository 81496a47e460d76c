#!/bin/bash
# usage: bash tools/run_multi_gpu.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR tools/dp_check.py 2>&1 | grep dp_check
timeout 600 $TR bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_n$N.err > gpurun_out/bench_n$N.json
echo "stdout lines: $(wc -l < gpurun_out/bench_n$N.json)"; python - <<PY
import json; d=json.load(open("gpurun_out/bench_n$N.json")); print("bench N=$N", round(d["value"],1), "Mpx/s  ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), d["clocks"])
PY
timeout 300 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null > gpurun_out/bench_ref_n$N.json; echo "reference stdout lines: $(wc -l < gpurun_out/bench_ref_n$N.json)"; cut -c1-120 gpurun_out/bench_ref_n$N.json
