timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -s -k "u64_training" 2>&1 | grep -E "passed|failed|^\{" | head -6
TPZ_TRAIN_SIMT=1 timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -s -k "u64_training" 2>&1 | grep -E "passed|failed|^\{" | head -6
