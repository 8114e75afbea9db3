"""Probe: do two independent MSMs on two streams (two contexts on one GPU) overlap their latency-bound phases?"""
import sys, time, threading
sys.path.insert(0, '.')
import numpy as np
import czk_b200
n = 1 << 20
rng = np.random.Generator(np.random.PCG64(1))
ctxs, bases, scs = [], [], []
for k in range(2):
    c = czk_b200.Context(0)
    b = c.bases_synthetic(1, 11 + k, n, 1024)
    b.precompute(0)
    sc = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    sc[:, 3] &= np.uint64((1 << 60) - 1)
    ctxs.append(c); bases.append(b); scs.append(c.vec_from(sc))
    c.msm_bases(b, scs[-1], montgomery=False)
reps = 6
def run(k, r):
    for _ in range(r):
        ctxs[k].msm_bases(bases[k], scs[k], montgomery=False)
t = time.perf_counter(); run(0, reps); run(1, reps); seq = time.perf_counter() - t
t = time.perf_counter()
th = [threading.Thread(target=run, args=(k, reps)) for k in range(2)]
[x.start() for x in th]; [x.join() for x in th]
par = time.perf_counter() - t
print(f"2 x {reps} MSMs of 2^20: sequential {seq*1e3/(2*reps):.2f} ms/MSM, two streams {par*1e3/(2*reps):.2f} ms/MSM")
