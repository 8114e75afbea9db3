# new NTT: correctness (each step under its own timeout; a hang must not eat the budget), then timing
timeout 60 python -c "
import sys; sys.path.insert(0,'.')
import czk_b200
from oracle import binding as o
ctx = czk_b200.Context(0)
v = o.random_fr_mont(1, 1<<12)
for inv in (False, True):
    for coset in (False, True):
        got = ctx._ntt_host(v, inv, coset); exp = o.ntt(v, inv, coset)
        print('2^12', inv, coset, bool((got == exp).all()), flush=True)
" 2>&1 | tail -6
timeout 400 python -m pytest tests/test_gpu_ntt.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python tools/ntt_once.py 21 2>&1 | tail -12
timeout 400 python -m pytest tests/test_gpu_groth16.py tests/test_gpu_large.py tests/test_gpu_plonk_leaves.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_ntt.json 2> gpurun_out/r2_bench_ntt.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_ntt.json').read())
print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['phases_ms'])
PY
