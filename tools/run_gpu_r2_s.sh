# Round 2, GPU call S: whole GPU suite on the current build; cfg4 timing; ncu --set full of the halo / first-layer training kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2s_gpu_tests.log 2>&1; tail -6 gpurun_out/r2s_gpu_tests.log | cut -c1-300
timeout 200 python bench.py --steps 3 --extras cfg4,cfg4bn --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); [print(k, v.get('ms_per_step'), v.get('kernel_launches_per_step')) for k,v in d['extra'].items()]"
TPZ_TRAIN_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 500 --launch-count 110 --csv --log-file gpurun_out/r2s_launches_train.csv python tools/bench_extra.py --workloads train > /dev/null 2>&1; tail -1 gpurun_out/r2s_launches_train.csv | cut -c1-160
TPZ_TRAIN_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"halo|first_tc_train" --launch-skip 60 --launch-count 14 -o gpurun_out/r2s_ncu_train_halo python tools/bench_extra.py --workloads train > gpurun_out/r2s_ncu.log 2>&1; tail -2 gpurun_out/r2s_ncu.log | cut -c1-200
ncu -i gpurun_out/r2s_ncu_train_halo.ncu-rep --page raw --csv > gpurun_out/r2s_ncu_train_halo_raw.csv 2>/dev/null; wc -c gpurun_out/r2s_ncu_train_halo_raw.csv
