# Round 2, GPU call F: ncu --set full of the tcgen05 training kernels on the cfg4-sized layer (stall reasons)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad_tc_kernel" -c 6 -f -o gpurun_out/r2f_ncu_train_tc python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "conv_tc_kernels and 256" -p no:cacheprovider > gpurun_out/r2f_ncu.log 2>&1; tail -3 gpurun_out/r2f_ncu.log | cut -c1-200
ncu -i gpurun_out/r2f_ncu_train_tc.ncu-rep --page raw --csv > gpurun_out/r2f_ncu_train_tc_raw.csv 2>/dev/null; ls -la gpurun_out/r2f_* | cut -c1-150
