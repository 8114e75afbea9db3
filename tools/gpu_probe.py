#!/usr/bin/env python3
"""First-contact probe for a GPU box: integer-pipe microbenchmarks and raw kernel timings."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np

import czk_b200
from oracle import binding as o


def main():
    ctx = czk_b200.Context(0)
    res = {}
    names = {0: "imad_wide", 1: "imad_lo_hi_pair", 2: "fr_mul", 3: "fq_mul", 4: "g1_madd"}
    for kind, name in names.items():
        for bps, thr in ((1, 128), (2, 128), (4, 128), (8, 128), (4, 256), (8, 256)):
            iters = 4000 if kind < 2 else (400 if kind < 4 else 60)
            ops, ms = ctx.microbench(kind, bps, thr, iters)
            res[f"{name}_b{bps}_t{thr}"] = ops
            print(f"{name:18s} blocks/SM={bps} threads={thr}: {ops:.3e} ops/s  ({ms:.3f} ms)", flush=True)
    import torch

    for log_d in (16, 20, 21, 22, 24):
        n = 1 << log_d
        v = ctx.vec_from(o.random_fr_mont(log_d, n))
        ctx.ntt_in_place(v, log_d)
        ctx.sync()
        t = time.perf_counter()
        reps = 5
        for _ in range(reps):
            ctx.ntt_in_place(v, log_d)
        ctx.sync()
        dt = (time.perf_counter() - t) / reps
        res[f"ntt_2^{log_d}_ms"] = dt * 1e3
        print(f"NTT 2^{log_d}: {dt*1e3:.3f} ms  ({64*n/dt/1e9:.1f} GB/s algorithmic)", flush=True)
        v.free()
    for curve, name in ((1, "g1"), (2, "g2")):
        for log_n in (16, 18, 20, 21):
            if curve == 2 and log_n > 20:
                continue
            n = 1 << log_n
            t = time.perf_counter()
            b = ctx.bases_synthetic(curve, 1, n, 1024)
            gen_s = time.perf_counter() - t
            sc = ctx.vec_from(o.random_fr_mont(log_n, n))
            ctx.msm_bases(b, sc)
            t = time.perf_counter()
            reps = 3
            for _ in range(reps):
                ctx.msm_bases(b, sc)
            dt = (time.perf_counter() - t) / reps
            res[f"msm_{name}_2^{log_n}_ms"] = dt * 1e3
            print(f"MSM {name} 2^{log_n}: {dt*1e3:.2f} ms (bases generated in {gen_s:.2f} s)", flush=True)
            b.free()
            sc.free()
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "probe.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
