# Round 2, GPU call L: warp-specialised halo fwd/dgrad + MN-major halo wgrad (probe, parity, A/B, launch list)
mkdir -p gpurun_out
timeout 120 python tools/debug_wgrad_halo.py > gpurun_out/r2l_wgrad_probe.log 2>&1; tail -30 gpurun_out/r2l_wgrad_probe.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -s -k "conv_tc_kernels" -p no:cacheprovider > gpurun_out/r2l_tc_kernels.log 2>&1; tail -3 gpurun_out/r2l_tc_kernels.log | cut -c1-300
TPZ_TRAIN_HALO_WGRAD=0 timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -k "conv_tc_kernels" -p no:cacheprovider 2>&1 | tail -2 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider -k "not conv_tc_kernels" 2>&1 | tail -4 | cut -c1-300
for wg in 0 1; do
echo "{\"TPZ_TRAIN_HALO_WGRAD\": $wg}"
TPZ_TRAIN_HALO_WGRAD=$wg timeout 200 python bench.py --steps 3 --extras cfg4,cfg4bn --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); [print(k, v.get('ms_per_step'), v.get('kernel_launches_per_step')) for k,v in d['extra'].items()]"
done
TPZ_TRAIN_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 500 --launch-count 100 --csv --log-file gpurun_out/r2l_launches_train_halo.csv python tools/bench_extra.py --workloads train > /dev/null 2>&1; tail -1 gpurun_out/r2l_launches_train_halo.csv | cut -c1-200
