mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
