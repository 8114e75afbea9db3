"""BASELINE.json config 5 - the standalone BLS12-377 G1 MSM + Fr NTT sweep 2^16 .. 2^24 on one GPU - with bit-exact parity
against the oracle at EVERY size that is timed (the oracle's multi-threaded Pippenger / radix-2 transform takes seconds at
these sizes on the GPU box's host cores).  Timings are written to gpurun_out/sweep_parity.json next to the verdicts; the
timing-only tool is tools/sweep.py.  Sizes follow SURVEY.md 8(d): bases with every 1024-th entry at infinity, uniform
scalars, G2 at 2^16 and 2^20 + 1."""
import json
import time
from pathlib import Path

import numpy as np
import pytest
from helpers import jac_to_affine_ints

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
RESULTS = {}


def _record(kind, key, value):
    RESULTS.setdefault(kind, {})[str(key)] = value
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "sweep_parity.json").write_text(json.dumps(RESULTS, indent=1))


def _time(fn, reps):
    fn()
    t = time.perf_counter()
    for _ in range(reps):
        r = fn()
    return (time.perf_counter() - t) / reps * 1e3, r


@pytest.mark.parametrize("n", [1 << 16, 1 << 18, 1 << 20, (1 << 21) - 1, 1 << 22, 1 << 24])
def test_sweep_msm_g1(ctx, oracle, n):
    G = oracle.G1
    b = ctx.bases_synthetic(1, 0x377 + n % 1009, n, 1024)
    try:
        xy, inf = b.numpy()
        sc = oracle.random_fr_mont(0x377 + n % 1013, n)
        dsc = ctx.vec_from(sc)
        out, isinf = G.msm(xy, inf, sc, threads=oracle.cpu_threads())
        exp = None if isinf else G.affine_to_ints(out)[0]
        ms, got = _time(lambda: ctx.msm_bases(b, dsc), 3 if n <= (1 << 22) else 1)
        assert jac_to_affine_ints(G, got) == exp, "windowed path"
        rec = {"windowed_ms": ms, "parity": "bit-exact vs oracle"}
        if n <= (1 << 22):  # the resident-table path the provers use (the 2^24 table would be 20 GB: timed by tools/sweep.py)
            b.precompute(0)
            ms_t, got = _time(lambda: ctx.msm_bases(b, dsc), 3)
            assert jac_to_affine_ints(G, got) == exp, "merged-window table path"
            rec["table_ms"] = ms_t
        _record("msm_g1", n, rec)
    finally:
        b.free()


@pytest.mark.parametrize("n", [1 << 16, (1 << 20) + 1])
def test_sweep_msm_g2(ctx, oracle, n):
    G = oracle.G2
    b = ctx.bases_synthetic(2, 0x99 + n % 1009, n, 1024)
    try:
        xy, inf = b.numpy()
        sc = oracle.random_fr_mont(0x99 + n % 1013, n)
        dsc = ctx.vec_from(sc)
        out, isinf = G.msm(xy, inf, sc, threads=oracle.cpu_threads())
        exp = None if isinf else G.affine_to_ints(out)[0]
        ms, got = _time(lambda: ctx.msm_bases(b, dsc), 2)
        assert jac_to_affine_ints(G, got) == exp
        b.precompute(0)
        ms_t, got = _time(lambda: ctx.msm_bases(b, dsc), 2)
        assert jac_to_affine_ints(G, got) == exp
        _record("msm_g2", n, {"windowed_ms": ms, "table_ms": ms_t, "parity": "bit-exact vs oracle"})
    finally:
        b.free()


@pytest.mark.parametrize("log_d", [16, 18, 20, 21, 22, 24])
def test_sweep_ntt(ctx, czk, oracle, log_d):
    n = 1 << log_d
    v = oracle.random_fr_mont(0x47 + log_d, n)
    th = oracle.cpu_threads()
    rec = {}
    for name, inv, cos in (("fft", False, False), ("ifft", True, False), ("coset_fft", False, True), ("coset_ifft", True, True)):
        dv = ctx.vec_from(v)
        ctx.ntt_in_place(dv, log_d, inv, cos)
        assert (dv.numpy() == oracle.ntt(v, inv, cos, threads=th)).all(), (log_d, name)
        ctx.sync()
        t = time.perf_counter()
        for _ in range(3):
            ctx.ntt_in_place(dv, log_d, inv, cos)
        ctx.sync()
        rec[name + "_ms"] = (time.perf_counter() - t) / 3 * 1e3
    rec["parity"] = "bit-exact vs oracle, all four transforms"
    _record("ntt", n, rec)
