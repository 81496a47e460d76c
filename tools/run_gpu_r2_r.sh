# Round 2, GPU call R: tensor-core first layer (fwd + wgrad) for the training net; suite; A/B; launch list
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -s -k "first_layer_tc" -p no:cacheprovider > gpurun_out/r2r_first.log 2>&1; grep -E "rel err|passed|failed|Error" gpurun_out/r2r_first.log | cut -c1-200 | head -20
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider -k "not first_layer_tc" > gpurun_out/r2r_train_tests.log 2>&1; tail -6 gpurun_out/r2r_train_tests.log | cut -c1-300
for f in 0 1; do
echo "{\"TPZ_TRAIN_FIRST_TC\": $f}"
TPZ_TRAIN_FIRST_TC=$f timeout 200 python bench.py --steps 3 --extras cfg4,cfg4bn --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); [print(k, v.get('ms_per_step'), v.get('kernel_launches_per_step')) for k,v in d['extra'].items()]"
done
TPZ_TRAIN_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 500 --launch-count 110 --csv --log-file gpurun_out/r2r_launches_train.csv python tools/bench_extra.py --workloads train > /dev/null 2>&1; tail -1 gpurun_out/r2r_launches_train.csv | cut -c1-200
# dense engine = model-level C ABI by default: ABI tests, parity tests, headline bench with the C-side dominant-kernel timing
timeout 600 python -m pytest tests/test_gpu_model_abi.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | cut -c1-300
timeout 300 python bench.py --steps 4 --warmup 3 --extras none --no-cpu-baseline 2>gpurun_out/r2r_bench.err > gpurun_out/r2r_bench.json; python -c "
import json; d=json.load(open('gpurun_out/r2r_bench.json')); print(round(d['value'],1), d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], d['roofline']['frac'], d['roofline']['launches_timed'])"; tail -2 gpurun_out/r2r_bench.err | cut -c1-300
