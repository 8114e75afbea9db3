nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multiparty.py -x -q -m gpu 2>&1 | tail -5
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_2gpu.json').read().strip().splitlines()[-1])
    print('N=2 ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['phases_ms'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2_bench_2gpu.err').read()[-3000:])
PY
