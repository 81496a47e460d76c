# Round 2, GPU call A: full GPU suite with the new parity tests (metrics printed), smoke, fast-split validation + A/B,
# headline bench, 3-D denoiser in auto vs fast precision.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2a_gpu_tests.log 2>&1; tail -15 gpurun_out/r2a_gpu_tests.log | cut -c1-300
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
TPZ_TRAIN_SPLIT=fast timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider > gpurun_out/r2a_train_tests_fast_split.log 2>&1; tail -3 gpurun_out/r2a_train_tests_fast_split.log | cut -c1-300
(echo '{"split": "default"}'; timeout 150 python tools/bench_extra.py --workloads train,train_bn; echo '{"split": "fast"}'; TPZ_TRAIN_SPLIT=fast timeout 150 python tools/bench_extra.py --workloads train,train_bn) 2>gpurun_out/r2a_bench_split_ab.err | tee gpurun_out/r2a_bench_split_ab.jsonl | cut -c1-260
timeout 300 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; cut -c1-600 gpurun_out/r2a_bench.json
(echo '{"precision": "auto"}'; timeout 200 python tools/bench_extra.py --workloads denoise3d,denoise --steps 3; echo '{"precision": "fast"}'; TPZ_PRECISION=fast timeout 200 python tools/bench_extra.py --workloads denoise3d --steps 3) 2>gpurun_out/r2a_bench_extra.err | tee gpurun_out/r2a_bench_extra.jsonl | cut -c1-260
