mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
for pr in 0 auto 0 auto; do
  if [ $pr = auto ]; then unset TPZ_TC_PAIR; else export TPZ_TC_PAIR=$pr; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_pair_$pr.err > gpurun_out/bench_pair_$pr.json
  python - <<PY
import json; d=json.load(open("gpurun_out/bench_pair_$pr.json")); print("pair=$pr", "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "dom ms", round(d["roofline"]["ms_per_launch"],2), d["clocks"])
PY
done
unset TPZ_TC_PAIR
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_pair.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 python tools/bench_extra.py --workloads denoise,denoise3d --steps 4 2>&1 | tail -3 | cut -c1-250
TPZ_TC_PAIR=0 timeout 600 python tools/bench_extra.py --workloads denoise,denoise3d --steps 4 2>&1 | tail -3 | cut -c1-250
