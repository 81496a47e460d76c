# Round 2, GPU call K: halo-resident training fwd/dgrad kernel (parity, A/B, launch list); model ABI + sampler re-check
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x -s -k "conv_tc_kernels" -p no:cacheprovider > gpurun_out/r2k_tc_kernels.log 2>&1; tail -3 gpurun_out/r2k_tc_kernels.log | cut -c1-300; grep -E "rel err" gpurun_out/r2k_tc_kernels.log | sort | uniq -c | sort -rn | head -5
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_model_abi.py tests/test_gpu_sampler.py -m gpu -q -p no:cacheprovider -k "not conv_tc_kernels" 2>&1 | tail -6 | cut -c1-300
for halo in 0 1; do
echo "{\"TPZ_TRAIN_HALO\": $halo}"
TPZ_TRAIN_HALO=$halo timeout 200 python bench.py --steps 3 --extras cfg4,cfg4bn --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); [print(k, v.get('ms_per_step'), v.get('kernel_launches_per_step')) for k,v in d['extra'].items()]"
done
TPZ_TRAIN_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 500 --launch-count 100 --csv --log-file gpurun_out/r2k_launches_train_halo.csv python tools/bench_extra.py --workloads train > /dev/null 2>&1; tail -1 gpurun_out/r2k_launches_train_halo.csv | cut -c1-200
