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
