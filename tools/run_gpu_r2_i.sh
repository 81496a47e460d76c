# Round 2, GPU call I (2 GPUs): model ABI re-check, sampler iterator, then data-parallel checks and the bench with extras at N=2
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model_abi.py tests/test_gpu_sampler.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR tools/dp_check.py 2>&1 | grep -E "dp_check|Error|error" | head -5
TPZ_TRAIN_GRAPH=dp timeout 300 $TR tools/dp_check.py 2>&1 | grep -E "dp_check|Error|error" | head -5
timeout 900 $TR bench.py --gpus 2 --steps 8 --warmup 3 2>gpurun_out/r2i_bench_n2.err > gpurun_out/r2i_bench_n2.json; tail -3 gpurun_out/r2i_bench_n2.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/r2i_bench_n2.json"))
print("bench N=2", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["clocks"])
for k,v in d.get("extra",{}).items():
    print(k, json.dumps(v)[:600])
PY
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
