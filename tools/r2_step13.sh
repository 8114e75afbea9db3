# mixed-radix NTT + Plonk wiring over the reference's wire domain: parity tests, bench line, reference arm
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_plonk.py -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_pytest_mixed.log; cat gpurun_out/r2_pytest_mixed.log
timeout 300 python bench.py --workload plonk --log-n 18 --steps 5 --warmup 3 > gpurun_out/r2_bench_plonk_1gpu.json 2> gpurun_out/r2_bench_plonk_1gpu.err; cut -c1-300 gpurun_out/r2_bench_plonk_1gpu.json; tail -3 gpurun_out/r2_bench_plonk_1gpu.err
timeout 600 python bench.py --workload plonk --log-n 18 --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_plonk_reference.json 2> gpurun_out/r2_bench_plonk_reference.err; cut -c1-300 gpurun_out/r2_bench_plonk_reference.json; tail -3 gpurun_out/r2_bench_plonk_reference.err
timeout 120 python tools/proof.py -p plonk -c squaring --computation-size 4096 local 2>&1 | tail -3
