# Round 2, GPU call X (8 GPUs): the driver's scaling step at N=8 -- default bench (cfg2 weak, cfg3 weak, cfg4 strong with NCCL, cfg5 strong) + dp_check
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR tools/dp_check.py 2>&1 | grep -E "dp_check|Error|error|capture" | head -5
( time timeout 900 $TR bench.py --gpus 8 --steps 8 --warmup 3 ) 2>gpurun_out/r2x_bench_n8.err > gpurun_out/r2x_bench_n8.json; tail -4 gpurun_out/r2x_bench_n8.err | cut -c1-200
python - <<PY
import json
d=json.load(open("gpurun_out/r2x_bench_n8.json"))
print("bench N=8", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"], "frac", d["roofline"]["frac"], d["clocks"])
for k,v in d.get("extra",{}).items():
    print(k, json.dumps(v)[:900])
PY
timeout 300 $TR bench.py --impl reference --gpus 8 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
