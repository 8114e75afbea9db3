timeout 900 python -m pytest tests/test_gpu_multiparty.py -x -q -m gpu 2>&1 | tail -5
grep -h "opens over" gpurun_out/mp_pytest_spdz_2*.log | cut -c1-300
for t in p2p nccl; do
  CZK_SHARE_TRANSPORT=$t timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/r2_bench_2gpu_$t.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$t', d['ms_per_step'], d['e2e']['value'], d['phases_ms']['witness_map'])"
done
