# end-of-round verification on one B200: GPU test-suite, smoke, headline bench, secondary workloads, same-GPU library baseline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; python - <<PY
import json; d=json.load(open("gpurun_out/bench_final.json")); print("bench", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3), d["clocks"], "cpu", round(d["cpu_baseline"]["value"],3), d["cpu_baseline"]["cores"])
PY
timeout 900 python tools/bench_extra.py --workloads denoise,train,train_tf32,train_bn,denoise3d --steps 6 2>/dev/null | tee gpurun_out/bench_extra_final.jsonl | cut -c1-200
# torch+cuDNN (TF32) on the same GPU: dense ResNet8-u64, U-Net patch, training step with and without BatchNorm
timeout 600 python tools/torch_cudnn_baseline.py 2>/dev/null | tee gpurun_out/torch_cudnn_same_gpu.jsonl | cut -c1-200
