# Plonk lines with the final binary: N = 4, 2 (on a 4-GPU box) 
mkdir -p gpurun_out
for N in 4 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --workload plonk --log-n 18 --steps 5 --warmup 3 > gpurun_out/r2_bench_plonk_${N}gpu.json 2> gpurun_out/r2_bench_plonk_${N}gpu.err; echo "plonk $N rc=$?"
done
timeout 300 python bench.py --workload plonk --log-n 18 --steps 5 --warmup 3 > gpurun_out/r2_bench_plonk_1gpu.json 2> gpurun_out/r2_bench_plonk_1gpu.err; echo "plonk 1 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29581 tests/mp_groth16_check.py --scheme spdz > gpurun_out/r2_mp_spdz_4.log 2>&1; echo "spdz rc=$?"; grep -o "parity ok" gpurun_out/r2_mp_spdz_4.log | wc -l
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_plonk_?gpu.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('phases_ms'))
    except Exception as e:
        print(f, 'failed', e)
PY
