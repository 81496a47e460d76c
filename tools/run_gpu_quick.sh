timeout 300 python tools/debug_train_u64.py 64 2>&1 | tail -35 | cut -c1-110
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -s 2>&1 | grep -E "passed|failed|^\{|worst" | cut -c1-400 | head
timeout 300 python tools/bench_extra.py --workloads train 2>/dev/null | cut -c1-200
