mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fused_first or resnet8_u64 or unet_pretrained" 2>&1 | tail -3 || exit 1
TPZ_FIRST=tc timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_stage.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
grep -E "first_tc" gpurun_out/launches_stage.csv | tail -2 | cut -c1-300
