# bit-slice reduction: adds per thread (32 / 16 / 8) for G1 2^21 and G2 2^20
for d in 32 16 8; do
  echo "div $d:"
  CZK_BITSUM_DIV=$d timeout 120 python tools/msm_once.py 1 21 0 2>&1 | grep -E "curve|rror"
  CZK_BITSUM_DIV=$d timeout 120 python tools/msm_once.py 2 20 0 2>&1 | grep -E "curve|rror"
done
CZK_BITSUM_DIV=8 timeout 200 python -m pytest tests/test_gpu_msm.py -x -q 2>&1 | tail -2
for d in 32 8; do CZK_BITSUM_DIV=$d timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('div $d ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'])
"; done
