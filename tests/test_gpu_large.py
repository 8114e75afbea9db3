"""Bit-exact parity at the BASELINE.json sizes, through the path the prover uses:
resident bases + merged-window table + the batched-affine bucket accumulation.
  * G1 MSM at 2^21 - 1 and 2^22 terms, G2 MSM at 2^20 + 1 terms vs the oracle's Pippenger
    (algebra/ec/src/msm/variable_base.rs:12-106 restated; identity of algebra/test-templates/src/msm.rs:16-33);
  * one whole SPDZ Groth16 proof at 2^20 constraints on a REAL device-generated CRS vs oracle.groth16_prove on the
    downloaded key (mpc-snarks/src/groth/prover.rs:66-177), also pairing-verified.
The oracle needs tens of seconds of CPU per case; everything else in the GPU tier stays at sizes it finishes in seconds."""
import numpy as np
import pytest
from helpers import jac_to_affine_ints

pytestmark = pytest.mark.gpu


def _msm_case(ctx, oracle, curve, n, seed):
    G = oracle.G1 if curve == 1 else oracle.G2
    b = ctx.bases_synthetic(curve, seed, n, 1024).precompute(0)
    try:
        xy, inf = b.numpy()
        sc = oracle.random_fr_mont(seed + 1, n)
        got = jac_to_affine_ints(G, ctx.msm_bases(b, ctx.vec_from(sc)))
        out, isinf = G.msm(xy, inf, sc, threads=oracle.cpu_threads())
        assert got == (None if isinf else G.affine_to_ints(out)[0])
    finally:
        b.free()


@pytest.mark.parametrize("n", [(1 << 21) - 1, 1 << 22])
def test_msm_g1_at_sweep_sizes_matches_oracle(ctx, oracle, n):
    _msm_case(ctx, oracle, 1, n, seed=0x377 + n % 97)


def test_msm_g2_2_20_plus_1_matches_oracle(ctx, oracle):
    _msm_case(ctx, oracle, 2, (1 << 20) + 1, seed=0x3770)


def test_spdz_proof_2_20_bit_exact_vs_oracle(ctx, czk, oracle):
    """BASELINE config C2's per-party work, one party: the benchmark's own key and path, every proof element compared."""
    ctx.net_init(0, 1, None)
    n_sq = 1 << 20
    rng = np.random.Generator(np.random.PCG64(0x377))
    toxic = rng.integers(0, 1 << 64, size=(7, 4), dtype=np.uint64)
    toxic[:, 3] &= np.uint64((1 << 60) - 1)
    dpk = czk.groth16_setup(ctx, n_sq, toxic)  # the key bench.py proves with
    try:
        pk = dpk.to_host()
        chain = oracle.squaring_chain(oracle.random_fr_mont(5, 1)[0], n_sq)
        r, s = oracle.random_fr_mont(6, 1), oracle.random_fr_mont(7, 1)
        got = czk.groth16_prove(ctx, czk.SCHEME_SPDZ, dpk, chain, r[0], s[0])
        exp = oracle.groth16_prove(oracle.SCHEME_SPDZ, n_sq, [chain], r, s, pk, threads=oracle.cpu_threads(), want_h=False)
        assert exp["ok"]
        assert (got["proof"] == exp["proof"]).all() and (got["proof_inf"] == exp["proof_inf"]).all()
        assert (got["proof_sh"] == exp["proof_sh"][0]).all()
        assert czk.groth16_verify(czk.pk_verifying_key(dpk), chain[n_sq:n_sq + 1], got["proof"], got["proof_inf"])
    finally:
        dpk.free()
