# final 8-GPU lines: SPDZ parity over the peer-memory transport, the Groth16 SPDZ 2^20 bench line, BASELINE config 4 (GSZ 2^22)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tests/mp_groth16_check.py --scheme spdz > gpurun_out/r2_mp_spdz_$N.log 2>&1; echo "spdz rc=$?"; grep -o "parity ok" gpurun_out/r2_mp_spdz_$N.log | wc -l; grep -o "opens over [a-zA-Z ()]*" gpurun_out/r2_mp_spdz_$N.log | sort | uniq -c
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench rc=$?"
if [ "$N" = "8" ]; then
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus $N --scheme gsz --log-n 22 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${N}gpu_gsz22.json 2> gpurun_out/r2_bench_${N}gpu_gsz22.err; echo "gsz rc=$?"
else
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus $N --workload plonk --log-n 18 --steps 5 --warmup 3 > gpurun_out/r2_bench_plonk_${N}gpu.json 2> gpurun_out/r2_bench_plonk_${N}gpu.err; echo "plonk rc=$?"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29564 tests/mp_groth16_check.py --scheme gsz > gpurun_out/r2_mp_gsz_$N.log 2>&1; echo "gsz check rc=$?"; grep -o "parity ok" gpurun_out/r2_mp_gsz_$N.log | wc -l
fi
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_*${N}gpu*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['config'].get('share_transport'), d.get('phases_ms'))
    except Exception as e:
        print(f, 'failed', e); print(open(f.replace('.json', '.err')).read()[-1500:])
PY
