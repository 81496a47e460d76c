timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -s -k "proj_wgrad" 2>&1 | grep -E "passed|failed|rel err|^E  *assert" | head -14
