# Round 2, GPU call Q (2 GPUs): data-parallel step captured in a CUDA graph by default; dp_check; cfg4 at N=2 graph vs eager
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR tools/dp_check.py 2>&1 | grep -E "dp_check|Error|error|capture" | head -5
for mode in 1 nodp; do
echo "{\"TPZ_TRAIN_GRAPH\": \"$mode\"}"
TPZ_TRAIN_GRAPH=$mode timeout 400 $TR bench.py --gpus 2 --steps 3 --warmup 3 --extras cfg4,cfg4bn --no-cpu-baseline 2>gpurun_out/r2q_bench_n2_$mode.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); [print(k, v.get('ms_per_step'), v.get('kernel_launches_per_step'), v.get('collectives')) for k,v in d['extra'].items()]"
grep -iE "capture|error" gpurun_out/r2q_bench_n2_$mode.err | head -3
done
