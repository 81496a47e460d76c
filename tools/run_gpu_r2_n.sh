# Round 2, GPU call N: MN-major (SWIZZLE_128B_BASE32B) probe + halo wgrad, scatter dgrad, split-K forward, classifier-head kernels
mkdir -p gpurun_out
timeout 120 python tools/lab_mn_major.py > gpurun_out/r2n_lab_mn_major.log 2>&1; grep -E "match|Error|error" gpurun_out/r2n_lab_mn_major.log | cut -c1-200
timeout 120 python tools/debug_wgrad_halo.py > gpurun_out/r2n_wgrad_probe.log 2>&1; grep -E "max rel" gpurun_out/r2n_wgrad_probe.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider > gpurun_out/r2n_train_tests.log 2>&1; tail -12 gpurun_out/r2n_train_tests.log | cut -c1-300
echo '{"all on"}'
timeout 200 python bench.py --steps 3 --extras cfg4,cfg4bn --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); [print(k, v.get('ms_per_step'), v.get('kernel_launches_per_step')) for k,v in d['extra'].items()]"
echo '{"halo wgrad off"}'
TPZ_TRAIN_HALO_WGRAD=0 timeout 200 python bench.py --steps 3 --extras cfg4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); [print(k, v.get('ms_per_step'), v.get('kernel_launches_per_step')) for k,v in d['extra'].items()]"
TPZ_TRAIN_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 500 --launch-count 110 --csv --log-file gpurun_out/r2n_launches_train.csv python tools/bench_extra.py --workloads train > /dev/null 2>&1; tail -1 gpurun_out/r2n_launches_train.csv | cut -c1-200
