# Round 2, GPU call E: 256-thread tcgen05 training kernels: parity, step time, launch lists (mma.sync vs tcgen05)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "conv_tc_kernels" -p no:cacheprovider 2>&1 | tail -3 | cut -c1-300
TPZ_TRAIN_TC=1 timeout 400 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider -k "not conv_tc_kernels" 2>&1 | tail -4 | cut -c1-300
for tc in 0 1; do
echo "{\"TPZ_TRAIN_TC\": $tc}"
TPZ_TRAIN_TC=$tc timeout 200 python bench.py --steps 3 --extras cfg4,cfg4bn --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); [print(k, v.get('ms_per_step'), v.get('kernel_launches_per_step')) for k,v in d['extra'].items()]"
TPZ_TRAIN_GRAPH=0 TPZ_TRAIN_TC=$tc timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 500 --launch-count 100 --csv --log-file gpurun_out/r2e_launches_train_tc$tc.csv python tools/bench_extra.py --workloads train > /dev/null 2>&1; tail -2 gpurun_out/r2e_launches_train_tc$tc.csv | cut -c1-200
done
