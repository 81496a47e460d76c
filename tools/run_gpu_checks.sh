mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python tools/bench_extra.py --workloads denoise --steps 6 2>&1 | tail -2 | cut -c1-220
TPZ_DENOISE_PIPELINE=0 timeout 600 python tools/bench_extra.py --workloads denoise --steps 6 2>&1 | tail -1 | cut -c1-220
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; python - <<PY
import json; d=json.load(open("gpurun_out/bench_final.json")); print("bench", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3), d["clocks"], d["cpu_baseline"]["value"])
PY
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
