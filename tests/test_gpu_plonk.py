"""GPU parity for the Plonk wiring argument (czk_plonk_prove_wiring, include/czk_plonk.h) against the oracle's restatement
of mpc-plonk/src/lib.rs:110-258,343-400 (oracle/czk_oracle_plonk.inc): same committer key, same wire / wiring polynomials,
same (stand-in) transcript => bit-identical commitments, challenges, opened values and opening proofs.  One party here;
the n-party run over NCCL is tests/mp_groth16_check.py (torchrun), which compares every rank's proof SHARES as well."""
import numpy as np
import pytest
from helpers import kzg_powers, plonk_wiring_instance

pytestmark = pytest.mark.gpu


def _prove_both(ctx, czk, oracle, scheme_name, log_d, powers, p, w, seed, precompute=False, mixed=False):
    scheme = {"plain": czk.SCHEME_PLAIN, "additive": czk.SCHEME_ADDITIVE, "spdz": czk.SCHEME_SPDZ}[scheme_name]
    oscheme = {"plain": oracle.SCHEME_PLAIN, "additive": oracle.SCHEME_ADDITIVE, "spdz": oracle.SCHEME_SPDZ}[scheme_name]
    exp = oracle.plonk_prove_wiring(oscheme, p[None], w, powers, seed=seed, threads=oracle.cpu_threads())
    assert exp["status"] == 1
    b = ctx.bases_upload(1, powers)
    if precompute:
        b.precompute(0)
    try:
        dp = ctx.vec_from(p)
        got = czk.plonk_prove_wiring(ctx, scheme, b, log_d, dp, ctx.vec_from(p) if scheme_name == "spdz" else None, ctx.vec_from(w), seed=seed,
                                     mixed=mixed)
    finally:
        b.free()
    return got, exp


@pytest.mark.parametrize("scheme_name", ["plain", "additive", "spdz"])
@pytest.mark.parametrize("log_d", [3, 10])
def test_plonk_wiring_matches_oracle_single_party(ctx, czk, oracle, scheme_name, log_d):
    ctx.net_init(0, 1, None)
    D = 1 << log_d
    powers = kzg_powers(D, 0x5eed + log_d)
    p, w = plonk_wiring_instance(log_d, seed=100 + log_d)
    got, exp = _prove_both(ctx, czk, oracle, scheme_name, log_d, powers, p, w, seed=7)
    for k in exp["proof"]:
        assert (got["proof"][k] == exp["proof"][k]).all(), k
    # one party: its share of every opening proof is the proof itself
    assert (got["proof_share"]["open_pf_xy"] == exp["share_pf_xy"][0]).all()
    assert (got["proof_share"]["open_pf_inf"] == exp["share_pf_inf"][0]).all()
    assert (np.array(oracle.fr_to_ints(got["proof"]["open_val"][2:3])) == 1).all()  # t(w^(k-1)) = 1: the instance is a valid wiring


def test_plonk_wiring_2_14_random_polynomials_with_table(ctx, czk, oracle):
    """A larger domain through the merged-window table path of the commitment / opening MSMs; random p and w (the argument's
    arithmetic does not need a satisfiable instance to be compared)."""
    ctx.net_init(0, 1, None)
    log_d = 14
    D = 1 << log_d
    g1, _ = oracle.generators()
    ks = oracle.random_fr_mont(3, 2)
    powers = oracle.G1.gen_progression(g1, ks[0], ks[1], D, threads=oracle.cpu_threads())  # any distinct points serve as a key here
    p, w = oracle.random_fr_mont(11, D), oracle.random_fr_mont(12, D)
    got, exp = _prove_both(ctx, czk, oracle, "spdz", log_d, powers, p, w, seed=99, precompute=True)
    for k in exp["proof"]:
        assert (got["proof"][k] == exp["proof"][k]).all(), k


@pytest.mark.parametrize("scheme_name", ["plain", "additive", "spdz"])
@pytest.mark.parametrize("log_m", [0, 2, 8])
def test_plonk_wiring_over_the_mixed_radix_wire_domain(ctx, czk, oracle, scheme_name, log_m):
    """The reference's circ.domains.wires is MixedRadixEvaluationDomain::new(3 n_gates) (relations/flat.rs:282-300):
    czk_plonk_prove_wiring_mixed over 3 * 2^log_m points against the oracle's mixed-radix restatement, valid instances."""
    ctx.net_init(0, 1, None)
    D = 3 << log_m
    powers = kzg_powers(D, 0xfeed + log_m)
    p, w = plonk_wiring_instance(None, seed=300 + log_m, size=D)
    got, exp = _prove_both(ctx, czk, oracle, scheme_name, log_m, powers, p, w, seed=17, mixed=True)
    for k in exp["proof"]:
        assert (got["proof"][k] == exp["proof"][k]).all(), k
    assert (got["proof_share"]["open_pf_xy"] == exp["share_pf_xy"][0]).all()
    assert (np.array(oracle.fr_to_ints(got["proof"]["open_val"][2:3])) == 1).all()


def test_plonk_wiring_mixed_3_2_12_with_table(ctx, czk, oracle):
    ctx.net_init(0, 1, None)
    log_m = 12
    D = 3 << log_m
    g1, _ = oracle.generators()
    ks = oracle.random_fr_mont(5, 2)
    powers = oracle.G1.gen_progression(g1, ks[0], ks[1], D, threads=oracle.cpu_threads())
    p, w = oracle.random_fr_mont(21, D), oracle.random_fr_mont(22, D)
    got, exp = _prove_both(ctx, czk, oracle, "spdz", log_m, powers, p, w, seed=5, precompute=True, mixed=True)
    for k in exp["proof"]:
        assert (got["proof"][k] == exp["proof"][k]).all(), k
