mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "unet or denoise or pool_upsample_last or fcnn" 2>&1 | tail -4
for m in auto tc; do
  echo "TPZ_LAST=$m"; TPZ_LAST=$m timeout 600 python tools/bench_extra.py --workloads denoise --steps 4 2>&1 | tail -2 | cut -c1-220
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_unet2d.csv python tools/unet_patch.py 2048 > /dev/null 2>&1
grep -E "conv_last_tiled" gpurun_out/launches_unet2d.csv | tail -1 | cut -c1-260
