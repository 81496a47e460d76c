mkdir -p gpurun_out
# 1. correctness of the CTA-pair mode (forced everywhere it is eligible)
TPZ_TC_PAIR=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -6
# 2. per-layer A/B
for pr in 0 1; do
TPZ_TC_PAIR=$pr timeout 300 python - <<PY 2>&1 | grep "^layer" | sed "s/^/pair=$pr /"
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
import layer_bench as L
for (cin, co, k, dil) in ((64, 64, 3, 2), (64, 64, 3, 4), (64,128,3,4), (128, 128, 3, 4), (128, 128, 3, 8), (128, 256, 5, 4), (96, 64, 5, 1), (48, 48, 3, 1)):
    L.layer(cin, co, k, dil, variant='v2')
L.layer(128, 128, 3, 8, variant='v2', residual_src=True)
PY
done
# 3. whole network
for pr in 0 1 auto; do
  if [ $pr = auto ]; then unset TPZ_TC_PAIR; else export TPZ_TC_PAIR=$pr; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_pair_$pr.json
  python - <<PY
import json; d=json.load(open("gpurun_out/bench_pair_$pr.json")); print("pair=$pr", "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "dom ms", round(d["roofline"]["ms_per_launch"],2), d["clocks"])
PY
done
