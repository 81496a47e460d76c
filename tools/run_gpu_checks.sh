mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python tools/bench_extra.py --workloads denoise3d 2>&1 | tail -1 | cut -c1-250
TPZ_DENOISE_GRAPH=0 timeout 600 python tools/bench_extra.py --workloads denoise3d 2>&1 | tail -1 | cut -c1-250
