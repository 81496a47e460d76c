# Round 2, GPU call C: new kernels (3-D Cout=1 tail, training-step CUDA graph) + the whole GPU suite; cfg4 / cfg5 numbers
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2c_gpu_tests.log 2>&1; tail -5 gpurun_out/r2c_gpu_tests.log | cut -c1-400
timeout 600 python bench.py --steps 4 --extras cfg4,cfg4bn,cfg5 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -3 gpurun_out/r2c_bench.err | cut -c1-300; python - <<PY
import json
d=json.load(open("gpurun_out/r2c_bench.json"))
for k,v in d.get("extra",{}).items():
    print(k, json.dumps(v)[:700])
PY
TPZ_PRECISION=fast timeout 300 python bench.py --steps 4 --extras cfg5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fast', json.dumps(d['extra'])[:600])"
