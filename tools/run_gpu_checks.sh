mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "unet or denoise or seeded or fcnn" 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_unet2d.csv python tools/unet_patch.py 2048 > /dev/null 2>&1
timeout 600 python tools/bench_extra.py --workloads denoise --steps 4 2>&1 | tail -2 | cut -c1-220
