# Round 2, GPU call AB: gather-GEMM prefetch moved behind the proxy fence (MEMBAR.ALL.CTA finding): parity, u32 / u64 step, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3 | cut -c1-300
timeout 300 python bench.py --steps 3 --extras cfg4,cfg4bn,cfg4u64 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); [print(k, v.get('ms_per_step'), v.get('kernel_launches_per_step')) for k,v in d['extra'].items()]"
TPZ_TRAIN_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 500 --launch-count 110 --csv --log-file gpurun_out/r2ab_launches_train.csv python tools/bench_extra.py --workloads train > /dev/null 2>&1; tail -1 gpurun_out/r2ab_launches_train.csv | cut -c1-100
