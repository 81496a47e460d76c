# BatchNorm training path on one B200: kernel/e2e parity tests first, then the step timing, then the whole GPU suite
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -s -k "batchnorm" > gpurun_out/bn_tests.log 2>&1; tail -25 gpurun_out/bn_tests.log | cut -c1-600
timeout 200 python tools/bench_extra.py --workloads train_bn,train 2>gpurun_out/bench_train_bn.err | tee gpurun_out/bench_train_bn.jsonl | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1; tail -8 gpurun_out/gpu_tests.log | cut -c1-400
