# Round 2, GPU call H: model-level C ABI (packed bytes vs the Python packer, bit-identical forward, plain-C host), whole suite
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model_abi.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2h_model_abi.log 2>&1; grep -E "first layer|^step|score_c|passed|failed|Error" gpurun_out/r2h_model_abi.log | cut -c1-200 | head -60
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_model_abi.py 2>&1 | tail -6 | cut -c1-300
