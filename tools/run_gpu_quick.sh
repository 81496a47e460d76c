timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -k "conv_mma" 2>&1 | grep -E "passed|failed|assert .*<|^E  *assert" | head -12
