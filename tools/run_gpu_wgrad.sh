# first-layer wgrad: row-strip kernel vs the previous kernel (A/B), parity tests of the training path
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q > gpurun_out/train_tests.log 2>&1; tail -12 gpurun_out/train_tests.log | cut -c1-400
(timeout 120 python tools/bench_extra.py --workloads train,train_tf32,train_bn; TPZ_FIRST_WGRAD=v1 timeout 120 python tools/bench_extra.py --workloads train,train_tf32) 2>gpurun_out/bench_wgrad.err | tee gpurun_out/bench_wgrad_ab.jsonl | cut -c1-260
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 80 --csv --log-file gpurun_out/launches_train_bn.csv python tools/bench_extra.py --workloads train_bn > /dev/null 2>&1; tail -3 gpurun_out/launches_train_bn.csv | cut -c1-200
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 50 --csv --log-file gpurun_out/launches_train.csv python tools/bench_extra.py --workloads train > /dev/null 2>&1; tail -3 gpurun_out/launches_train.csv | cut -c1-200
