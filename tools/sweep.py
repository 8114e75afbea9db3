#!/usr/bin/env python3
"""BASELINE config 5: standalone BLS12-377 G1 MSM + Fr NTT sweep 2^16..2^24 on one GPU.
Writes gpurun_out/sweep.json; inputs are resident (device-generated bases, uploaded scalars)."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np

import czk_b200


def rand_fr(rng, n):
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)  # < 2^252 < r: valid canonical / Montgomery limbs
    return a


def ref_adds(n):
    lg = n.bit_length() - 1 if n & (n - 1) == 0 else n.bit_length()
    c = 3 if n < 32 else lg * 69 // 100 + 2
    w = (253 + c - 1) // c
    return n * w + 2 * ((1 << c) - 1) * w + 253


def main():
    ctx = czk_b200.Context(0)
    rng = np.random.Generator(np.random.PCG64(0x377))
    peak, _ = ctx.microbench(0, 8, 256, 2000)
    out = {"imad_wide_peak_per_s": peak, "msm_g1": {}, "msm_g1_table": {}, "msm_g2": {}, "ntt": {}}
    for log_n in (16, 18, 20, 21, 22, 24):
        n = (1 << log_n) - (1 if log_n == 21 else 0)
        b = ctx.bases_synthetic(1, 0x377 + log_n, n, 1024)
        sc = ctx.vec_from(rand_fr(rng, n))
        for _ in range(3):  # warm-up: module loading, workspace growth, clocks
            ctx.msm_bases(b, sc)
        reps = 5 if log_n <= 21 else 2
        ctx.msm_stats(1, reset=True)
        t = time.perf_counter()
        for _ in range(reps):
            ctx.msm_bases(b, sc)
        dt = (time.perf_counter() - t) / reps
        st = ctx.msm_stats(1)
        out["msm_g1"][str(n)] = {"ms": dt * 1e3, "device_ms": st["msm_ms"] / reps, "accumulate_ms": st["accumulate_ms"] / reps,
                                 "adds_ref_per_s": ref_adds(n) / dt}
        print(f"MSM G1 n={n}: {dt*1e3:.2f} ms  ({ref_adds(n)/dt:.3e} reference-adds/s)", flush=True)
        # the same MSM over a resident base set with its merged-window table (how the prover runs its CRS queries)
        t = time.perf_counter()
        b.precompute(0)
        ctx.sync()
        pre_s = time.perf_counter() - t
        ctx.msm_bases(b, sc)
        ctx.msm_stats(1, reset=True)
        t = time.perf_counter()
        for _ in range(reps):
            ctx.msm_bases(b, sc)
        dt = (time.perf_counter() - t) / reps
        st = ctx.msm_stats(1)
        out["msm_g1_table"][str(n)] = {"ms": dt * 1e3, "device_ms": st["msm_ms"] / reps, "accumulate_ms": st["accumulate_ms"] / reps,
                                       "adds_ref_per_s": ref_adds(n) / dt, "table_build_s": pre_s}
        print(f"MSM G1 n={n} (precomputed table): {dt*1e3:.2f} ms  ({ref_adds(n)/dt:.3e} reference-adds/s; table built in {pre_s:.2f} s)", flush=True)
        b.free()
        sc.free()
    for log_n in (16, 20):
        n = (1 << log_n) + (1 if log_n == 20 else 0)
        b = ctx.bases_synthetic(2, 0x99 + log_n, n, 1024)
        sc = ctx.vec_from(rand_fr(rng, n))
        ctx.msm_bases(b, sc)
        t = time.perf_counter()
        for _ in range(3):
            ctx.msm_bases(b, sc)
        dt = (time.perf_counter() - t) / 3
        out["msm_g2"][str(n)] = {"ms": dt * 1e3}
        print(f"MSM G2 n={n}: {dt*1e3:.2f} ms", flush=True)
        b.free()
        sc.free()
    for log_d in (16, 18, 20, 21, 22, 24):
        n = 1 << log_d
        v = ctx.vec_from(rand_fr(rng, n))
        res = {}
        for name, inv, cos in (("fft", 0, 0), ("ifft", 1, 0), ("coset_fft", 0, 1), ("coset_ifft", 1, 1)):
            ctx.ntt_in_place(v, log_d, bool(inv), bool(cos))
            ctx.sync()
            t = time.perf_counter()
            for _ in range(5):
                ctx.ntt_in_place(v, log_d, bool(inv), bool(cos))
            ctx.sync()
            dt = (time.perf_counter() - t) / 5
            res[name] = {"ms": dt * 1e3, "algorithmic_GBps": 64 * n / dt / 1e9}
        out["ntt"][str(n)] = res
        print(f"NTT 2^{log_d}: " + ", ".join(f"{k} {v_['ms']:.3f} ms" for k, v_ in res.items()), flush=True)
        v.free()
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "sweep.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
