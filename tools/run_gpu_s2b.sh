# session-2 check: training-path tests (PReLU extractors, epoch round trip), then ncu --set full captures of the U-Net
# patch kernels and of every layer of the ResNet8-u64 4096^2 step (raw metric pages exported to CSV)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_train.py -m gpu -q -k "activation or prelu or round_trip or batchnorm" > gpurun_out/train_tests_b.log 2>&1; tail -6 gpurun_out/train_tests_b.log | cut -c1-600
timeout 75 ncu --set full --clock-control none -k regex:"tc_conv2|first_tc|conv_last_tiled" -s 64 -c 32 -f -o gpurun_out/ncu_unet python tools/unet_patch.py > gpurun_out/ncu_unet.log 2>&1
ncu -i gpurun_out/ncu_unet.ncu-rep --page raw --csv > gpurun_out/ncu_unet_raw.csv 2>/dev/null; ls -la gpurun_out/ncu_unet* | cut -c1-120
timeout 110 ncu --set full --clock-control none -k regex:"tc_conv2|first_tc" -s 24 -c 8 -f -o gpurun_out/ncu_resnet python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_resnet.log 2>&1
ncu -i gpurun_out/ncu_resnet.ncu-rep --page raw --csv > gpurun_out/ncu_resnet_raw.csv 2>/dev/null; ls -la gpurun_out/ncu_resnet* | cut -c1-120
rm -f gpurun_out/ncu_unet.ncu-rep gpurun_out/ncu_resnet.ncu-rep
