mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_plonk.py -x -q 2>&1 | tail -2
for s in spdz additive; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tests/mp_groth16_check.py --scheme $s > gpurun_out/r2_mp_${s}_2.log 2>&1; echo "$s rc=$?"; grep -o "parity ok" gpurun_out/r2_mp_${s}_2.log | wc -l; done
timeout 300 python bench.py --workload plonk --log-n 18 --steps 5 --warmup 3 > gpurun_out/r2_bench_plonk_1gpu.json 2> gpurun_out/r2_bench_plonk_1gpu.err; echo "plonk 1 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 2 --workload plonk --log-n 18 --steps 5 --warmup 3 > gpurun_out/r2_bench_plonk_2gpu.json 2> gpurun_out/r2_bench_plonk_2gpu.err; echo "plonk 2 rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_plonk_[12]gpu.json')):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('phases_ms'))
PY
