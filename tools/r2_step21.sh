mkdir -p gpurun_out
N=4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/r2_bench_4gpu.json').read().strip().splitlines()[-1])
print('N=4 ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['config'].get('share_transport'), d['phases_ms'])
PY
