"""GPU parity for the GSZ20 (honest-majority Shamir) path: czk_gsz_* and czk_groth16_prove(scheme GSZ) against the
oracle's restatement of mpc-algebra/src/share/gsz20/mod.rs.  One party here (1 GPU, t = 0: the share domain has one
point); the 4- and 8-party runs with real Shamir shares and failing degree checks are tests/mp_groth16_check.py
--scheme gsz under torchrun, one rank per GPU."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(oracle, pymodel, n_sq, seed):
    rnd = random.Random(seed)
    toxic = [rnd.randrange(1, pymodel.R_MOD) for _ in range(7)]
    pk = oracle.groth16_setup(n_sq, oracle.fr_from_ints(toxic), threads=oracle.cpu_threads())
    chain = oracle.squaring_chain(oracle.fr_from_ints([rnd.randrange(pymodel.R_MOD)])[0], n_sq)
    r = oracle.fr_from_ints([rnd.randrange(pymodel.R_MOD)])
    s = oracle.fr_from_ints([rnd.randrange(pymodel.R_MOD)])
    return pk, chain, r, s


@pytest.mark.parametrize("n_sq", [10, 13, 1 << 10])  # 13: D = 16 with odd lengths inside ip_check's halving
def test_gsz_prove_matches_oracle_single_party(ctx, czk, oracle, pymodel, n_sq):
    ctx.net_init(0, 1, None)
    pk, chain, r, s = _setup(oracle, pymodel, n_sq, seed=100 + n_sq)
    exp = oracle.groth16_prove_gsz(1, n_sq, chain, r[0], s[0], pk, threads=oracle.cpu_threads())
    assert exp["ok"]
    dpk = czk.ProvingKey.upload(ctx, pk)
    h = czk.groth16_witness_map(ctx, czk.SCHEME_GSZ, n_sq, chain)
    assert (h == exp["h"]).all()
    ctx.gsz_check_products()  # drain the triple the witness map queued
    before = ctx.gsz_stats()
    got = czk.groth16_prove(ctx, czk.SCHEME_GSZ, dpk, chain, r[0], s[0])
    assert (got["proof"] == exp["proof"]).all() and (got["proof_inf"] == exp["proof_inf"]).all()
    assert (got["proof_sh"] == exp["proof"]).all()  # every party's share is the value itself under the stubs
    assert (got["field_check"] == exp["field_check"]).all()
    assert (got["group_check_x"] == exp["group_check_x"]).all()
    assert (got["group_check_yz"] == exp["group_check_yz"]).all() and (got["group_check_inf"] == exp["group_check_inf"]).all()
    # the same protocol steps: king computations and opens
    assert got["king_computes"] - before["king_computes"] == exp["king_computes"]
    assert got["opens"] - before["opens"] == exp["opens"]
    dpk.free()


def test_gsz_batch_mul_and_check(ctx, czk, oracle):
    ctx.net_init(0, 1, None)
    k = 1000
    x, y = oracle.random_fr_mont(21, k), oracle.random_fr_mont(22, k)
    xs, ys = ctx.vec_from(x), ctx.vec_from(y)
    ctx.gsz_batch_mul(xs, ys)
    assert (xs.numpy() == oracle.fr_mul(x, y)).all()
    fx, fy, fz = ctx.gsz_check_products()
    assert (oracle.fr_mul(fx[None, :], fy[None, :])[0] == fz).all()
    # opening at any degree with one party returns the share
    assert (ctx.gsz_open(xs, 0).numpy() == oracle.fr_mul(x, y)).all()


def test_gsz_at_baseline_domain_2_21(ctx, czk, oracle):
    """The Hadamard product of config C4's shape at D = 2^21 (one GPU): product, queued triple, full ip_check."""
    ctx.net_init(0, 1, None)
    k = 1 << 21
    x, y = oracle.random_fr_mont(31, k), oracle.random_fr_mont(32, k)
    xs, ys = ctx.vec_from(x), ctx.vec_from(y)
    ctx.gsz_batch_mul(xs, ys)
    assert (xs.numpy() == oracle.fr_mul(x, y)).all()
    fx, fy, fz = ctx.gsz_check_products()
    assert (oracle.fr_mul(fx[None, :], fy[None, :])[0] == fz).all()


# ---------------------------------------------------------------------------------------------------------------------
# open_degree_vec (gsz20/mod.rs:440-459) on genuine N-party Shamir shares, on ONE GPU: czk_diag_gsz_open_gathered feeds
# k_gsz_open - the kernel behind czk_gsz_open / czk_gsz_king_compute - a party-major matrix of gathered shares.
def _shamir_shares(oracle, parties, coeffs):
    """coeffs: (t + 1, k, 4) Montgomery coefficients of k polynomials -> (parties, k, 4): party j holds p_i(w^j).
    The powers w^(j c) come from the oracle's own sharing of the monomial X^c (orc_gsz_share)."""
    tp1, k = coeffs.shape[0], coeffs.shape[1]
    one = oracle.fr_from_ints([1])[0]
    out = np.zeros((parties, k, 4), np.uint64)
    for c in range(tp1):
        mono = np.zeros((c + 1, 4), np.uint64)
        mono[c] = one
        w_pows = oracle.gsz_share(parties, mono)  # (parties, 4): w^(j c)
        for j in range(parties):
            out[j] = oracle.fr_add(out[j], oracle.fr_mul(coeffs[c], np.broadcast_to(w_pows[j], (k, 4)).copy()))
    return out


@pytest.mark.parametrize("parties,k", [(2, 1 << 10), (3, 1000), (4, 1 << 10), (8, 1 << 10), (6, 333), (8, 1 << 21)])
def test_gsz_open_degree_vec_on_n_party_shares(ctx, czk, oracle, parties, k):
    t = (parties - 1) // 2
    coeffs = np.stack([oracle.random_fr_mont(500 + 10 * parties + c, k) for c in range(t + 1)])
    shares = _shamir_shares(oracle, parties, coeffs)
    # spot-check the construction against the oracle's per-polynomial sharing and opening
    for i in (0, k - 1):
        assert (oracle.gsz_share(parties, coeffs[:, i]) == shares[:, i]).all()
        val, ok = oracle.gsz_open(shares[:, i], t)
        assert ok and (val == coeffs[0, i]).all()
    gathered = ctx.vec_from(shares.reshape(parties * k, 4))
    out, flag = ctx.gsz_open_gathered(gathered, parties, t, k)
    assert flag == 0 and (out.numpy() == coeffs[0]).all()
    if t >= 1:
        # the product of two degree-t sharings has degree 2t: it opens at 2t and must FAIL the degree-t check
        prod = oracle.fr_mul(shares.reshape(-1, 4), shares.reshape(-1, 4))
        secrets2 = oracle.fr_mul(coeffs[0], coeffs[0])
        out, flag = ctx.gsz_open_gathered(ctx.vec_from(prod), parties, 2 * t, k)
        assert flag == 0 and (out.numpy() == secrets2).all()
        out, flag = ctx.gsz_open_gathered(ctx.vec_from(prod), parties, t, k)
        assert flag == 1, "degree check did not fire"
        assert (out.numpy() == secrets2).all()  # the value at 0 is still the interpolated one (the reference asserts after computing it)
        _, ok = oracle.gsz_open(prod.reshape(parties, k, 4)[:, 0], t)
        assert not ok
    # one corrupted share among many: the flag must fire (n > t + 1 parties) wherever it sits
    if parties >= 3:
        bad = shares.copy()
        bad[parties - 1, k // 3] = oracle.random_fr_mont(999, 1)[0]
        _, flag = ctx.gsz_open_gathered(ctx.vec_from(bad.reshape(parties * k, 4)), parties, t, k)
        assert flag == 1
