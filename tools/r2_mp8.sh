N=${1:-8}
nvidia-smi -L | wc -l
for scheme in spdz gsz; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tests/mp_groth16_check.py --scheme $scheme > gpurun_out/r2_mp_${scheme}_$N.log 2>&1; echo "$scheme rc=$?"; grep -c "parity ok" gpurun_out/r2_mp_${scheme}_$N.log
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().splitlines()[-1])
    print('N=$N ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['phases_ms'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2_bench_${N}gpu.err').read()[-2000:])
PY
