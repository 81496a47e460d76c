# Round 2, GPU call AA: launch list of the resnet8_u64 training step (cfg4 secondary)
mkdir -p gpurun_out
TPZ_TRAIN_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 400 --launch-count 120 --csv --log-file gpurun_out/r2aa_launches_train_u64.csv python tools/bench_extra.py --workloads train_u64 > /dev/null 2>&1; tail -1 gpurun_out/r2aa_launches_train_u64.csv | cut -c1-120
