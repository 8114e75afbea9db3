import sys
sys.path.insert(0, '.')
import numpy as np
import czk_b200
curve = int(sys.argv[1]) if len(sys.argv) > 1 else 1
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
pre = int(sys.argv[3]) if len(sys.argv) > 3 else -1   # -1: windowed; 0: merged auto; c: merged with window c
ctx = czk_b200.Context(0)
n = 1 << log_n
b = ctx.bases_synthetic(curve, 1, n, 1024)
if pre >= 0:
    b.precompute(pre)
rng = np.random.Generator(np.random.PCG64(1))
sc = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
sc[:, 3] &= np.uint64((1 << 60) - 1)
dsc = ctx.vec_from(sc)
ctx.msm_bases(b, dsc, montgomery=False)  # warm-up: module loading, workspace allocation
ctx.msm_stats(curve, reset=True)
for _ in range(3):
    ctx.msm_bases(b, dsc, montgomery=False)
st = ctx.msm_stats(curve)
print(f"curve {curve} n=2^{log_n} pre={pre}: accumulate {st['accumulate_ms']/3:.2f} ms, msm device {st['msm_ms']/3:.2f} ms")
