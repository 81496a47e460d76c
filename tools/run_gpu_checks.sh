mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fused_first or resnet8_u64" 2>&1 | tail -3 || exit 1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
for v in tc tc; do
  TPZ_FIRST=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_first_$v.json
  python - <<PY
import json; d=json.load(open("gpurun_out/bench_first_$v.json")); print("first=$v", "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "dom ms", round(d["roofline"]["ms_per_launch"],2), d["clocks"])
PY
done
TPZ_FIRST=tc timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_wide.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 python tools/bench_extra.py --workloads denoise,denoise3d --steps 4 2>&1 | tail -3 | cut -c1-250
