# First GPU call of the next round (one B200, ~3 min): validates the candidates left unmeasured at the end of round 1.
#  1. TPZ_TRAIN_SPLIT=fast (3-instruction TF32 operand split, profiles/r01_sass_train_mma_s2.md): training parity suite, then an
#     A/B of the training step against the default split on the same box
#  2. torch+cuDNN baseline on the same GPU incl. the BatchNorm training step
mkdir -p gpurun_out
TPZ_TRAIN_SPLIT=fast timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q > gpurun_out/train_tests_fast_split.log 2>&1; tail -4 gpurun_out/train_tests_fast_split.log | cut -c1-400
(echo '{"split": "default"}'; timeout 150 python tools/bench_extra.py --workloads train,train_bn; echo '{"split": "fast"}'; TPZ_TRAIN_SPLIT=fast timeout 150 python tools/bench_extra.py --workloads train,train_bn) 2>gpurun_out/bench_split_ab.err | tee gpurun_out/bench_split_ab.jsonl | cut -c1-240
timeout 400 python tools/torch_cudnn_baseline.py 2>/dev/null | tee gpurun_out/torch_cudnn_same_gpu.jsonl | cut -c1-200
