mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_pytest_gpu.log
timeout 60 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' 2>&1 | tail -2 > gpurun_out/r2_smoke.log; cat gpurun_out/r2_smoke.log
