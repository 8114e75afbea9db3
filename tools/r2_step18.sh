timeout 400 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_groth16.py -x -q 2>&1 | tail -2
for v in 1 0; do CZK_NTT_DIRECT_SCALE=$v timeout 120 python tools/ntt_once.py 21 2>&1 | grep "ifft+coset_fft  x6"; done
for v in 1 0 1 0; do CZK_NTT_DIRECT_SCALE=$v timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('direct $v ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'wm', d['phases_ms']['witness_map'])
"; done
