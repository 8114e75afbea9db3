timeout 400 python -m pytest tests/test_gpu_plonk.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --workload plonk --log-n 18 --steps 5 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('plonk ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['phases_ms'])
"
