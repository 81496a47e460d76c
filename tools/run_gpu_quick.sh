mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_unet3d.csv python tools/unet_patch.py 3d > /dev/null 2>&1
timeout 600 python tools/bench_extra.py --workloads denoise3d 2>/dev/null | cut -c1-220
