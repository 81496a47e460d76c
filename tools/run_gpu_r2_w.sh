# Round 2, GPU call W: halo wgrad on 32x32 blocks of wider layers (64->64, 128->128); training suite; step time; launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider > gpurun_out/r2w_train_tests.log 2>&1; tail -5 gpurun_out/r2w_train_tests.log | cut -c1-300
timeout 200 python bench.py --steps 3 --extras cfg4,cfg4bn --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); [print(k, v.get('ms_per_step'), v.get('kernel_launches_per_step')) for k,v in d['extra'].items()]"
TPZ_TRAIN_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 500 --launch-count 110 --csv --log-file gpurun_out/r2w_launches_train.csv python tools/bench_extra.py --workloads train > /dev/null 2>&1; tail -1 gpurun_out/r2w_launches_train.csv | cut -c1-120
