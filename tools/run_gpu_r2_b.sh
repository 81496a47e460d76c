# Round 2, GPU call B: bench.py with the extra block (cfg3/4/5 + same-GPU cuDNN baseline); launch lists of one 192^3 patch of
# the 3-D denoiser in auto and fast precision.
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -3 gpurun_out/r2b_bench.err | cut -c1-300; python - <<PY
import json
d=json.load(open("gpurun_out/r2b_bench.json"))
print("bench", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3), d["clocks"])
for k,v in d.get("extra",{}).items():
    print(k, json.dumps(v)[:900])
PY
for mode in auto fast; do
TPZ_PRECISION=$mode timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_unet3d_$mode.csv python tools/unet_patch.py 3d > gpurun_out/r2b_unet3d_$mode.log 2>&1; tail -2 gpurun_out/r2b_unet3d_$mode.log | cut -c1-200
done
