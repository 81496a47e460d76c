mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "unet or denoise or fcnn" 2>&1 | tail -3
timeout 600 python tools/bench_extra.py --workloads denoise --steps 6 2>&1 | tail -3 | cut -c1-250
TPZ_DENOISE_PIPELINE=0 timeout 600 python tools/bench_extra.py --workloads denoise --steps 6 2>&1 | tail -1 | cut -c1-250
TPZ_DENOISE_GRAPH=0 TPZ_DENOISE_PIPELINE=0 timeout 600 python tools/bench_extra.py --workloads denoise --steps 6 2>&1 | tail -1 | cut -c1-250
