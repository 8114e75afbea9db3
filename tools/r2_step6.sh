timeout 600 python -m pytest tests/test_gpu_plonk.py tests/test_gpu_plonk_leaves.py -x -q -m gpu 2>&1 | tail -4
timeout 200 python bench.py --workload plonk --log-n 18 --steps 5 --warmup 3 > gpurun_out/r2_bench_plonk_1gpu.json 2> gpurun_out/r2_bench_plonk_1gpu.err; tail -c 900 gpurun_out/r2_bench_plonk_1gpu.json; tail -3 gpurun_out/r2_bench_plonk_1gpu.err
timeout 120 python tools/proof.py -p plonk -c squaring --computation-size 1000 mpc --alg spdz 2>&1 | tail -3
