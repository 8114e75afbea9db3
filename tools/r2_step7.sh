timeout 500 python -m pytest tests/test_gpu_msm.py tests/test_golden_vectors.py tests/test_gpu_large.py -x -q -m gpu 2>&1 | tail -2
for args in "1 21" "1 20" "1 18" "2 20"; do timeout 120 python tools/msm_once.py $args 0 2>&1 | grep -E "curve|rror"; done
echo "== priorities on"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['value'], d['phases_ms'], d['g1_msm'])"
echo "== priorities off"; CZK_STREAM_PRIORITY=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['value'])"
timeout 200 python bench.py --workload plonk --log-n 18 --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('plonk', d['ms_per_step'], d['phases_ms'])"
