# Round 2, GPU call AF: ncu launch list of the headline bench command (model-level C ABI engine), per B200_PROFILING.md
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2af_launches_bench.csv python bench.py --steps 2 --warmup 1 --extras none --no-cpu-baseline > gpurun_out/r2af_bench_under_ncu.log 2>&1; tail -1 gpurun_out/r2af_launches_bench.csv | cut -c1-120; wc -l gpurun_out/r2af_launches_bench.csv
