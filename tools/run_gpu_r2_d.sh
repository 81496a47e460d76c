# Round 2, GPU call D: tcgen05 training kernels: kernel-level parity, then the training suite and the step time with them on
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -x -s -k "conv_tc_kernels" -p no:cacheprovider > gpurun_out/r2d_tc_kernels.log 2>&1; tail -25 gpurun_out/r2d_tc_kernels.log | cut -c1-300
TPZ_TRAIN_TC=1 timeout 400 python -m pytest tests/test_gpu_train.py tests/test_gpu_parity_r2.py -m gpu -q -p no:cacheprovider -k "not conv_tc_kernels" > gpurun_out/r2d_train_tests_tc.log 2>&1; tail -8 gpurun_out/r2d_train_tests_tc.log | cut -c1-300
for tc in 0 1; do for fl in 2 0; do
echo "{\"TPZ_TRAIN_TC\": $tc, \"TPZ_TRAIN_FLUSH\": $fl}"
TPZ_TRAIN_TC=$tc TPZ_TRAIN_FLUSH=$fl timeout 200 python bench.py --steps 3 --extras cfg4,cfg4bn --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); [print(k, v.get('ms_per_step'), v.get('kernel_launches_per_step'), v.get('last_out')) for k,v in d['extra'].items()]"
done; done
