#!/usr/bin/env python3
"""BASELINE config 3 shape (Plonk, SPDZ, 2^18 constraints): device time of the leaves one `prove_unit_product` +
KZG commit/open round reaches (mpc-plonk/src/lib.rs:110-190): 3 coset FFTs + 1 coset iFFT, partial_products,
batch_div, KZG commit (MSM 2^18) and KZG open (division by X - z + MSM).  Writes gpurun_out/plonk_leaves.json.
    python tools/plonk_leaves.py [log_n]          (torchrun --nproc-per-node P for P parties)"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np

import czk_b200
from czk_b200 import launch


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
    party = launch.Party()
    ctx, rank, world = party.ctx, party.rank, party.world
    scheme = czk_b200.SCHEME_SPDZ
    n = 1 << log_n
    rng = np.random.Generator(np.random.PCG64(7))
    vals = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    vals[:, 3] &= np.uint64((1 << 60) - 1)
    mine = launch.king_share_scatter(vals if rank == 0 else None, n, seed=3)
    powers = ctx.bases_synthetic(1, 0x18, n, 0).precompute(0)  # powers_of_g stand-in, resident with its table
    z = np.array([0x1111, 0x2222, 0x3333, 0x0444], np.uint64)
    out = {}

    def timed(name, fn, reps=5):
        fn()
        ctx.sync()
        launch.barrier()
        t = time.perf_counter()
        for _ in range(reps):
            fn()
        ctx.sync()
        out[name] = launch.max_over_ranks((time.perf_counter() - t) / reps * 1e3)

    x, xm = ctx.vec_from(mine), ctx.vec_from(mine)
    y, ym = ctx.vec_from(mine), ctx.vec_from(mine)
    timed("coset_fft_sh+mac", lambda: (ctx.ntt_in_place(x, log_n, False, True), ctx.ntt_in_place(xm, log_n, False, True)))
    timed("coset_ifft_sh+mac", lambda: (ctx.ntt_in_place(x, log_n, True, True), ctx.ntt_in_place(xm, log_n, True, True)))

    def fresh():
        x.upload(mine), xm.upload(mine), y.upload(mine), ym.upload(mine)

    def pp():
        fresh()
        ctx.share_partial_products(scheme, x, xm)

    def bd():
        fresh()
        ctx.share_batch_div(scheme, x, xm, y, ym)

    def bm():
        fresh()
        ctx.batch_mul(scheme, x, xm, y, ym)

    timed("upload_4_vectors", fresh)
    timed("batch_mul(+upload)", bm)
    timed("partial_products(+upload)", pp)
    timed("batch_div(+upload)", bd)
    fresh()
    timed("kzg_commit_msm", lambda: ctx.msm_bases(powers, x))
    timed("kzg_open(div+msm)", lambda: ctx.kzg_open(powers, x, z))
    timed("poly_div_linear", lambda: ctx.poly_div_linear(x, z))
    if rank == 0:
        res = {"workload": f"plonk leaves, spdz, n=2^{log_n}, {world} parties", "ms": out}
        print(json.dumps(res))
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / f"plonk_leaves_{world}p.json").write_text(json.dumps(res, indent=1))
    party.close()


if __name__ == "__main__":
    main()
