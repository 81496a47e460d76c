# Round 2, GPU call M: MN-major operand probe; sub-lattice strided dgrad (parity + A/B)
mkdir -p gpurun_out
timeout 120 python tools/lab_mn_major.py > gpurun_out/r2m_lab_mn_major.log 2>&1; tail -70 gpurun_out/r2m_lab_mn_major.log | cut -c1-250
TPZ_TRAIN_HALO_WGRAD=0 timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | cut -c1-300
for lat in 0 1; do
echo "{\"TPZ_TRAIN_LATTICE_DGRAD\": $lat}"
TPZ_TRAIN_HALO_WGRAD=0 TPZ_TRAIN_LATTICE_DGRAD=$lat timeout 200 python bench.py --steps 3 --extras cfg4,cfg4bn --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); [print(k, v.get('ms_per_step'), v.get('kernel_launches_per_step')) for k,v in d['extra'].items()]"
done
