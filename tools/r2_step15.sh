timeout 400 python -m pytest tests/test_gpu_groth16.py tests/test_gpu_gsz.py tests/test_gpu_setup.py -x -q 2>&1 | tail -2
for v in 1 0; do CZK_AB_SHARE_PLAN=$v timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('share $v ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['phases_ms'], d['config'].get('proof_verified'))
"; done
