mkdir -p gpurun_out
TPZ_RESIDUAL=epilogue timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "resnet or seeded or ragged" 2>&1 | tail -3
for rs in mma epilogue mma epilogue; do
  TPZ_RESIDUAL=$rs timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_res_$rs.json
  python - <<PY
import json; d=json.load(open("gpurun_out/bench_res_$rs.json")); print("residual=$rs", "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "dom ms", round(d["roofline"]["ms_per_launch"],2), d["clocks"])
PY
done
