# Round 2, GPU call V: what the driver runs at round end -- pytest -m gpu, smoke(), reference arm, default bench (timed wall clock)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > gpurun_out/r2v_gpu_tests.log 2>&1; tail -8 gpurun_out/r2v_gpu_tests.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r2v_bench_ref.json 2> gpurun_out/r2v_bench_ref.err; tail -4 gpurun_out/r2v_bench_ref.err | cut -c1-200; cut -c1-300 gpurun_out/r2v_bench_ref.json
( time timeout 1200 python bench.py ) > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; tail -4 gpurun_out/r2v_bench.err | cut -c1-200
python - <<PY
import json
d=json.load(open("gpurun_out/r2v_bench.json"))
print("bench N=1", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"], "frac", d["roofline"]["frac"], d["clocks"])
print("cpu_baseline", d.get("cpu_baseline"))
for k,v in d.get("extra",{}).items():
    print(k, json.dumps(v)[:700])
PY
