"""CPU tier: the oracle's restatement of the Plonk / KZG10 leaves (oracle/czk_oracle_plonk.inc) against algorithm-independent
identities - the only pins the reference offers for them (its debug_asserts in mpc-plonk/src/lib.rs:122-176 and the KZG
check equation poly-commit/src/kzg10/mod.rs:282-300)."""
import random

import numpy as np
import pytest


def _total(oracle, sh):
    tot = sh[0]
    for p in range(1, sh.shape[0]):
        tot = oracle.fr_add(tot, sh[p])
    return tot


@pytest.mark.parametrize("n", [1, 2, 9, 100])
def test_poly_div_linear_is_exact_division(oracle, pymodel, n):
    R = pymodel.R_MOD
    rnd = random.Random(n)
    coef = [rnd.randrange(R) for _ in range(n)]
    for z in (rnd.randrange(R), 0, 1, R - 1):
        q, rem = oracle.poly_div_linear(oracle.fr_from_ints(coef), oracle.fr_from_ints([z])[0])
        qi, ri = oracle.fr_to_ints(q), oracle.fr_to_ints(rem[None, :])[0]
        chk = [0] * n  # q * (X - z) + rem
        for i, c in enumerate(qi):
            chk[i + 1] = (chk[i + 1] + c) % R
            chk[i] = (chk[i] - c * z) % R
        chk[0] = (chk[0] + ri) % R
        assert chk == coef
        assert ri == sum(c * pow(z, i, R) for i, c in enumerate(coef)) % R  # remainder = p(z)


@pytest.mark.parametrize("parties", [1, 2, 3])
@pytest.mark.parametrize("scheme_name", ["additive", "spdz"])
def test_share_protocols_reconstruct(oracle, pymodel, parties, scheme_name):
    R = pymodel.R_MOD
    rnd = random.Random(7 * parties)
    scheme = oracle.SCHEME_SPDZ if scheme_name == "spdz" else oracle.SCHEME_ADDITIVE
    k = 33
    xv = [rnd.randrange(1, R) for _ in range(k)]
    yv = [rnd.randrange(1, R) for _ in range(k)]
    xsh = np.stack(oracle.king_share_batch(oracle.fr_from_ints(xv), parties, seed=1))
    ysh = np.stack(oracle.king_share_batch(oracle.fr_from_ints(yv), parties, seed=2))
    mac = (lambda a: a.copy()) if scheme_name == "spdz" else (lambda a: None)
    st, out, outm = oracle.share_op(oracle.SHARE_BATCH_INV, scheme, xsh, mac(xsh))
    assert st == 1 and oracle.fr_to_ints(_total(oracle, out)) == [pow(v, -1, R) for v in xv]
    if outm is not None:  # MAC key stubbed to 1: the MAC shares reconstruct to the same value
        assert oracle.fr_to_ints(_total(oracle, outm)) == [pow(v, -1, R) for v in xv]
    st, out, _ = oracle.share_op(oracle.SHARE_BATCH_DIV, scheme, xsh, mac(xsh), ysh, mac(ysh))
    assert st == 1 and oracle.fr_to_ints(_total(oracle, out)) == [a * pow(b, -1, R) % R for a, b in zip(xv, yv)]
    st, out, _ = oracle.share_op(oracle.SHARE_PARTIAL_PRODUCTS, scheme, xsh, mac(xsh))
    acc, exp = 1, []
    for v in xv:
        acc = acc * v % R
        exp.append(acc)
    assert st == 1 and oracle.fr_to_ints(_total(oracle, out)) == exp
    # a zero divisor: the reference's .inverse().unwrap() panics; the restatement reports it
    xv0 = list(xv)
    xv0[5] = 0
    z0 = np.stack(oracle.king_share_batch(oracle.fr_from_ints(xv0), parties, seed=3))
    st, _, _ = oracle.share_op(oracle.SHARE_BATCH_INV, scheme, z0, mac(z0))
    assert st == 0
    # a corrupted SPDZ MAC share must be caught by the open inside the protocol
    if scheme_name == "spdz" and parties > 1:
        bad = xsh.copy()
        bad[1, 0, 0] = bad[1, 0, 0] ^ np.uint64(1)
        st, _, _ = oracle.share_op(oracle.SHARE_BATCH_INV, scheme, xsh, bad)
        assert st == -1


def test_kzg_open_satisfies_the_check_equation(oracle, pymodel):
    R = pymodel.R_MOD
    G = oracle.G1
    n, tau = 24, 0xabcdef
    g1, _ = oracle.generators()
    powers = np.stack([G.scalar_mul(g1, oracle.fr_from_ints([pow(tau, i, R)])[0])[0] for i in range(n)])
    p = oracle.random_fr_mont(5, n)
    z = oracle.random_fr_mont(6, 1)[0]
    w, winf, ev = oracle.kzg_open(powers, None, p, z)
    assert not winf and (ev == oracle.poly_eval(p, z)).all()
    commit, cinf = G.msm(powers, None, p)
    gen = G.affine_to_ints(g1[None, :])[0]
    ci, wi = G.affine_to_ints(commit)[0], G.affine_to_ints(w[None, :])[0]
    evi, zi = oracle.fr_to_ints(ev[None, :])[0], oracle.fr_to_ints(z[None, :])[0]
    # e(C - v G, H) = e(w, (tau - z) H)  <=>  C - v G == (tau - z) w
    assert pymodel.g1_add(ci, pymodel.g1_neg(pymodel.g1_mul(gen, evi))) == pymodel.g1_mul(wi, (tau - zi) % R)
