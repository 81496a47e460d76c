mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python - <<'PY' 2>&1 | tail -20
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
import layer_bench as L
for (cin, co, k, dil) in ((64, 64, 3, 2), (64, 64, 3, 4), (64,128,3,4), (128, 128, 3, 4), (128, 128, 3, 8), (64, 64, 1, 1), (128, 256, 5, 4)):
    L.layer(cin, co, k, dil, variant='v2')
L.layer(64, 64, 3, 4, variant='v2', residual_src=True)
L.layer(128, 128, 3, 8, variant='v2', residual_src=True)
PY
for nt in 2 1; do TPZ_CO256_NTILE=$nt timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_nt$nt.json; python - <<PY
import json; d=json.load(open("gpurun_out/bench_nt$nt.json")); print("ntile $nt", round(d["value"],1), round(d["ms_per_step"],2), round(d["roofline"]["ms_per_launch"],2), d["clocks"])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 python tools/bench_extra.py --workloads denoise 2>&1 | tail -2
