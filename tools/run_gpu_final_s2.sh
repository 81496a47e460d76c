# end-of-session verification: the whole GPU suite, then smoke()
mkdir -p gpurun_out
timeout 85 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests_final_s2.log 2>&1; tail -5 gpurun_out/gpu_tests_final_s2.log | cut -c1-500
timeout 30 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/smoke_final_s2.log
