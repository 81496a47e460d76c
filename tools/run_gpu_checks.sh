# A/B checks of the fused first layer and the pipelined denoise host path on ONE box (box-to-box host speed varies)
mkdir -p gpurun_out
for v in tc im2col tc im2col; do
  TPZ_FIRST=$v timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_first_$v.json
  python - <<PY
import json; d=json.load(open("gpurun_out/bench_first_$v.json")); print("first=$v", "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "dom ms", round(d["roofline"]["ms_per_launch"],2), d["clocks"])
PY
done
for v in tc im2col; do
  TPZ_FIRST=$v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_first_$v.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
done
for p in 1 0 1 0; do
  echo "denoise pipeline=$p"; TPZ_DENOISE_PIPELINE=$p timeout 600 python tools/bench_extra.py --workloads denoise --steps 4 2>&1 | tail -1 | cut -c1-260
done
