# Round 2, GPU call AC (2 GPUs): final build -- cfg4 strong + weak under data parallelism (CUDA graph with NCCL), dp_check
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 200 $TR tools/dp_check.py 2>&1 | grep -E "dp_check|Error|error|capture" | head -3
timeout 400 $TR bench.py --gpus 2 --steps 4 --warmup 3 --extras cfg4 --no-cpu-baseline 2>gpurun_out/r2ac_bench_n2.err > gpurun_out/r2ac_bench_n2.json; python - <<PY
import json
d=json.load(open("gpurun_out/r2ac_bench_n2.json"))
print("N=2", round(d["value"],1), round(d["e2e"]["value"],1))
for k,v in d["extra"].items():
    print(k, v if isinstance(v,str) else {a: v.get(a) for a in ("ms_per_step","value","scaling","collectives")})
PY
grep -iE "capture|Error" gpurun_out/r2ac_bench_n2.err | head -3
