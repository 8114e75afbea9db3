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


# ---------------------------------------------------------------------------------------------------------------------
# The wiring argument end to end (oracle/czk_oracle_plonk.inc: orc_plonk_prove_wiring).  The reference holds no Plonk
# vectors, so the restatement is pinned by what the reference's own (commented-out) debug assertions and its verifier
# check: the opened values satisfy the wiring identities (mpc-plonk/src/lib.rs:161-189, 246-249 and the verifier at
# :460-560), every opening proof is the KZG10 witness of its commitment (poly-commit/src/kzg10/mod.rs:265-290, checked
# here in the exponent with the known tau), and an n-party run reveals exactly the single-prover proof.
def _g1_add(oracle, a_xy, a_inf, b_xy, b_inf):
    one = oracle.fr_from_ints([1])[0]
    xy = np.stack([a_xy, b_xy])
    inf = np.array([a_inf, b_inf], np.uint8)
    out, isinf = oracle.G1.msm(xy, inf, np.stack([one, one]))
    return (None if isinf else oracle.G1.affine_to_ints(out)[0])


def _g1_mul(oracle, xy, inf, k_int):
    out, isinf = oracle.G1.scalar_mul(xy, oracle.fr_from_ints([k_int])[0], inf)
    return out, isinf


@pytest.mark.parametrize("D", [8, 64, 12, 48])
def test_plonk_wiring_proof_identities_and_kzg(oracle, pymodel, D):
    """Power-of-two domains and the reference's own wire domain shape, 3 * 2^k (MixedRadixEvaluationDomain)."""
    from helpers import kzg_powers, plonk_wiring_instance

    R = pymodel.R_MOD
    tau = 0x1234567 + D
    powers = kzg_powers(D, tau)
    p, w = plonk_wiring_instance(None, seed=5 + D, size=D)
    res = oracle.plonk_prove_wiring(oracle.SCHEME_PLAIN, p[None], w, powers, seed=77)
    assert res["status"] == 1
    pf = res["proof"]
    y, z, r, x = oracle.fr_to_ints(pf["challenges"])
    v = oracle.fr_to_ints(pf["open_val"])
    t_wr, t_r, t_wk, f_wr, q_r, l2q_x, w_x, l1_x, p_x = v
    Z = lambda u: (pow(u, D, R) - 1) % R
    assert t_wk == 1                                                       # lib.rs:188  t(w^(k-1)) = 1
    assert (t_wr - t_r * f_wr - Z(r) * q_r) % R == 0                       # lib.rs:184-187
    assert ((p_x + y * x + z) * l1_x - (p_x + y * w_x + z) - l2q_x * Z(x)) % R == 0  # lib.rs:246-249
    # KZG10 openings in the exponent: (tau - point) * W == C - value * G
    dp = oracle.mixed_domain_params(D)
    omega, omega_inv = oracle.fr_to_ints([dp[0], dp[1]])
    g1, _ = oracle.generators()
    one = oracle.fr_from_ints([1])[0]
    p_cmt, p_cmt_inf = oracle.G1.msm(powers, None, p)
    w_cmt, w_cmt_inf = oracle.G1.msm(powers, None, w)
    cmts = {"l1": (pf["cmt_xy"][0], pf["cmt_inf"][0]), "t": (pf["cmt_xy"][1], pf["cmt_inf"][1]), "q": (pf["cmt_xy"][2], pf["cmt_inf"][2]),
            "l2q": (pf["cmt_xy"][3], pf["cmt_inf"][3]), "p": (p_cmt, p_cmt_inf), "w": (w_cmt, w_cmt_inf)}
    opens = [("t", omega * r % R), ("t", r), ("t", omega_inv), ("l1", omega * r % R), ("q", r), ("l2q", x), ("w", x), ("l1", x), ("p", x)]
    for slot, (name, point) in enumerate(opens):
        c_xy, c_inf = cmts[name]
        lhs, lhs_inf = _g1_mul(oracle, pf["open_pf_xy"][slot], int(pf["open_pf_inf"][slot]), (tau - point) % R)
        neg_vg, nv_inf = _g1_mul(oracle, g1, 0, (-v[slot]) % R)
        rhs = _g1_add(oracle, c_xy, int(c_inf), neg_vg, int(nv_inf))
        assert (None if lhs_inf else oracle.G1.affine_to_ints(lhs)[0]) == rhs, (slot, name)


@pytest.mark.parametrize("scheme_name,parties,D", [("additive", 2, 32), ("spdz", 2, 32), ("spdz", 3, 32), ("additive", 4, 32), ("spdz", 2, 24)])
def test_plonk_wiring_n_party_reveals_the_single_prover_proof(oracle, scheme_name, parties, D):
    from helpers import kzg_powers, plonk_wiring_instance

    powers = kzg_powers(D, 0xabcdef)
    p, w = plonk_wiring_instance(None, seed=21, size=D)
    single = oracle.plonk_prove_wiring(oracle.SCHEME_PLAIN, p[None], w, powers, seed=3)
    shares = oracle.king_share_batch(p, parties, seed=9)
    scheme = oracle.SCHEME_SPDZ if scheme_name == "spdz" else oracle.SCHEME_ADDITIVE
    multi = oracle.plonk_prove_wiring(scheme, shares, w, powers, seed=3)
    assert single["status"] == 1 and multi["status"] == 1
    for k in single["proof"]:
        assert (single["proof"][k] == multi["proof"][k]).all(), k
