# Round 2, GPU call T: C-packer range guard + first-layer bias gradient tests, accumulator-flush A/B on the halo kernels, HBM write bandwidth probe
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model_abi.py tests/test_gpu_parity_r2.py tests/test_gpu_train.py -m gpu -q -p no:cacheprovider -k "model or range_guard or first_layer or ge_binomial or golden" 2>&1 | tail -5 | cut -c1-300
timeout 120 python tools/membw_probe.py 2>&1 | tail -5
for fl in 2 0; do
echo "{\"TPZ_TRAIN_FLUSH\": $fl}"
TPZ_TRAIN_FLUSH=$fl timeout 200 python bench.py --steps 3 --extras cfg4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); [print(k, v.get('ms_per_step'), v.get('kernel_launches_per_step')) for k,v in d['extra'].items()]"
done
TPZ_TRAIN_FLUSH=0 TPZ_TRAIN_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 500 --launch-count 110 --csv --log-file gpurun_out/r2t_launches_train_noflush.csv python tools/bench_extra.py --workloads train > /dev/null 2>&1; tail -1 gpurun_out/r2t_launches_train_noflush.csv | cut -c1-120
