timeout 600 python -m pytest tests -m gpu -q -k "fused_first or unet or denoise" 2>&1 | tail -3
timeout 600 python tools/bench_extra.py --workloads denoise --steps 6 2>/dev/null | cut -c1-200
