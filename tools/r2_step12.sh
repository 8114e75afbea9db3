# bucket reduction: two-level grid sums vs the bit-slice kernel
for m in grid bits; do
  echo "reduce $m:"
  CZK_MSM_REDUCE=$m timeout 120 python tools/msm_once.py 1 21 0 2>&1 | grep -E "curve|rror"
  CZK_MSM_REDUCE=$m timeout 120 python tools/msm_once.py 2 20 0 2>&1 | grep -E "curve|rror"
  CZK_MSM_REDUCE=$m timeout 120 python tools/msm_once.py 1 18 0 2>&1 | grep -E "curve|rror"
done
timeout 300 python -m pytest tests/test_gpu_msm.py tests/test_gpu_groth16.py -x -q 2>&1 | tail -2
for m in grid bits; do CZK_MSM_REDUCE=$m timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('reduce $m ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'])
"; done
