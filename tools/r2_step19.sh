mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_groth16.py tests/test_gpu_ntt.py::test_mixed_radix_ntt_host_entry -x -q 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_multiparty.py -x -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_2gpu_b.json 2> gpurun_out/r2_bench_2gpu_b.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/r2_bench_2gpu_b.json').read().strip().splitlines()[-1])
print('N=2 ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['phases_ms'])
PY
