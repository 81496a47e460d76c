# Round 2, GPU call J: tcgen05 training kernels with three accumulators; model ABI diagnostics
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "conv_tc_kernels" -p no:cacheprovider 2>&1 | tail -3 | cut -c1-300
timeout 400 python -m pytest tests/test_gpu_train.py tests/test_gpu_parity_r2.py -m gpu -q -p no:cacheprovider -k "not conv_tc_kernels" 2>&1 | tail -4 | cut -c1-300
for tc in 0 1; do
echo "{\"TPZ_TRAIN_TC\": $tc}"
TPZ_TRAIN_TC=$tc timeout 200 python bench.py --steps 3 --extras cfg4,cfg4bn --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); [print(k, v.get('ms_per_step'), v.get('kernel_launches_per_step')) for k,v in d['extra'].items()]"
done
TPZ_TRAIN_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 500 --launch-count 100 --csv --log-file gpurun_out/r2j_launches_train_tc1.csv python tools/bench_extra.py --workloads train > /dev/null 2>&1; tail -1 gpurun_out/r2j_launches_train_tc1.csv | cut -c1-200
timeout 300 python -m pytest tests/test_gpu_model_abi.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2j_model_abi.log 2>&1; grep -E "python plans|AssertionError|passed|failed" gpurun_out/r2j_model_abi.log | cut -c1-250 | head -30
