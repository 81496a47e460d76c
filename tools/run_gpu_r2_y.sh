# Round 2, GPU call Y: Denoise3D with device-side statistics and pinned result (parity + cfg5 timing); ncu --set full of the remaining gather-GEMM launches
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -m gpu -q -p no:cacheprovider -k "3d or unet or Denoise or denoise" 2>&1 | tail -4 | cut -c1-300
timeout 300 python bench.py --steps 3 --extras cfg5,cfg3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); [print(k, {a: v.get(a) for a in ('value','ms_total','ms_per_patch','ms_per_image','e2e')}) for k,v in d['extra'].items()]"
TPZ_TRAIN_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad_tc_kernel" --launch-skip 40 --launch-count 11 -o gpurun_out/r2y_ncu_train_gather python tools/bench_extra.py --workloads train > gpurun_out/r2y_ncu.log 2>&1; tail -1 gpurun_out/r2y_ncu.log | cut -c1-200
ncu -i gpurun_out/r2y_ncu_train_gather.ncu-rep --page raw --csv > gpurun_out/r2y_ncu_train_gather_raw.csv 2>/dev/null; wc -c gpurun_out/r2y_ncu_train_gather_raw.csv
