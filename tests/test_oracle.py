"""Pin the C oracle (oracle/czk_oracle.c) against the independent Python big-int model and against the
identities the reference's own tests use."""
import random

import numpy as np
import pytest


def test_field_axioms_vs_bigint(oracle, pymodel):
    # curves/bls12_377/src/fields/tests.rs:177-238 style, against Python ints
    rnd = random.Random(1)
    for mod, nl, frm, to, mul, add, sub, inv in (
        (pymodel.R_MOD, 4, oracle.fr_from_ints, oracle.fr_to_ints, oracle.fr_mul, oracle.fr_add, oracle.fr_sub, oracle.fr_inv),
        (pymodel.Q_MOD, 6, oracle.fq_from_ints, oracle.fq_to_ints, oracle.fq_mul, oracle.fq_add, oracle.fq_sub, oracle.fq_inv),
    ):
        a = [rnd.randrange(mod) for _ in range(300)] + [0, 1, mod - 1, mod - 1]
        b = [rnd.randrange(mod) for _ in range(300)] + [mod - 1, mod - 1, mod - 1, 1]
        A, B = frm(a), frm(b)
        assert to(mul(A, B)) == [x * y % mod for x, y in zip(a, b)]
        assert to(add(A, B)) == [(x + y) % mod for x, y in zip(a, b)]
        assert to(sub(A, B)) == [(x - y) % mod for x, y in zip(a, b)]
        nz = [x or 1 for x in a]
        assert to(inv(frm(nz))) == [pow(x, -1, mod) for x in nz]


def test_fq2_vs_bigint(oracle, pymodel):
    rnd = random.Random(2)
    vals = [rnd.randrange(pymodel.Q_MOD) for _ in range(400)]
    A = oracle.fq_from_ints(vals[:200]).reshape(-1, 12)
    B = oracle.fq_from_ints(vals[200:]).reshape(-1, 12)
    pa = list(zip(vals[0:200:2], vals[1:200:2]))
    pb = list(zip(vals[200:400:2], vals[201:400:2]))
    prod = oracle.fq_to_ints(oracle.fq2_mul(A, B).reshape(-1, 6))
    sq = oracle.fq_to_ints(oracle.fq2_sqr(A).reshape(-1, 6))
    iv = oracle.fq_to_ints(oracle.fq2_inv(A).reshape(-1, 6))
    for i, (x, y) in enumerate(zip(pa, pb)):
        assert (prod[2 * i], prod[2 * i + 1]) == pymodel.fq2_mul(x, y)
        assert (sq[2 * i], sq[2 * i + 1]) == pymodel.fq2_mul(x, x)
        assert (iv[2 * i], iv[2 * i + 1]) == pymodel.fq2_inv(x)


@pytest.mark.parametrize("g", ["g1", "g2"])
def test_msm_pippenger_equals_naive_and_bigint_model(oracle, pymodel, g):
    """algebra/test-templates/src/msm.rs:16-33 (Pippenger == naive, compared as affine), plus an
    independent affine big-int evaluation."""
    G = oracle.G1 if g == "g1" else oracle.G2
    gen = pymodel.G1_GEN if g == "g1" else pymodel.G2_GEN
    mul = pymodel.g1_mul if g == "g1" else pymodel.g2_mul
    naive = pymodel.g1_msm_naive if g == "g1" else pymodel.g2_msm_naive
    rnd = random.Random(3)
    n = 48
    pts = [mul(gen, rnd.randrange(1, 10**6)) for _ in range(n)]
    pts[5] = None  # infinity base
    sc = [rnd.randrange(pymodel.R_MOD) for _ in range(n)]
    sc[3], sc[4], sc[7] = 0, 1, 1  # zero scalar (filtered), unit scalars (window-0 shortcut)
    xy, inf = G.affine_from_ints(pts)
    S = oracle.fr_from_ints(sc)
    r1, i1 = G.msm(xy, inf, S)
    r2, i2 = G.msm_naive(xy, inf, S)
    expect = naive(pts, sc)
    assert not i1 and not i2
    assert G.affine_to_ints(r1)[0] == G.affine_to_ints(r2)[0] == expect
    # window-parallel variant (the dormant Rayon split, variable_base.rs:36) is bit-identical
    r3, _ = G.msm(xy, inf, S, threads=4)
    assert (r3 == r1).all()
    # larger: 2^10 random points as in the template, Pippenger vs naive in C
    n = 1 << 10 if g == "g1" else 1 << 8
    from helpers import make_points

    xy = make_points(G, n, seed=12)
    S = oracle.random_fr_mont(13, n)
    assert (G.msm(xy, None, S, threads=4)[0] == G.msm_naive(xy, None, S)[0]).all()


def test_group_law_special_cases(oracle, pymodel):
    """short_weierstrass_jacobian.rs:594-596 (P+P), :598 (P + -P), infinity handling."""
    G = oracle.G1
    P = pymodel.g1_mul(pymodel.G1_GEN, 7)
    xy, _ = G.affine_from_ints([P, P, pymodel.g1_neg(P), None])
    one = oracle.fr_from_ints([1, 1, 1, 1])
    out, isinf = G.msm_naive(xy, np.array([0, 0, 0, 1], np.uint8), one)
    assert not isinf and G.affine_to_ints(out)[0] == P
    out, isinf = G.msm(xy[:2], None, one[:2])
    assert G.affine_to_ints(out)[0] == pymodel.g1_mul(P, 2)
    out, isinf = G.msm(xy[1:3], None, one[:2])
    assert isinf


@pytest.mark.parametrize("log_d", range(0, 9))
def test_ntt_vs_bigint_dft_and_serial_radix2(oracle, pymodel, log_d):
    """radix2/mod.rs:320-360 (FFT == evaluation), :381-491 (== the test-suite's serial CLRS FFT)."""
    n = 1 << log_d
    rnd = random.Random(log_d)
    v = [rnd.randrange(pymodel.R_MOD) for _ in range(n)]
    V = oracle.fr_from_ints(v)
    for inverse in (False, True):
        for coset in (False, True):
            got = oracle.fr_to_ints(oracle.ntt(V, inverse, coset))
            assert got == pymodel.ntt(v, inverse, coset, slow=(log_d <= 5)), (inverse, coset)
            assert (oracle.ntt(V, inverse, coset, threads=4) == oracle.ntt(V, inverse, coset)).all()
    assert (oracle.serial_radix2_fft(V) == oracle.ntt(V)).all()
    assert (oracle.serial_radix2_fft(V, True) == oracle.ntt(V, True)).all()


def test_ntt_roundtrip_and_horner_2_12(oracle, pymodel):
    n = 1 << 12
    V = oracle.random_fr_mont(5, n)
    assert (oracle.ntt(oracle.ntt(V), True) == V).all()
    assert (oracle.ntt(oracle.ntt(V, False, True), True, True) == V).all()
    d = oracle.domain_params(n)
    ev = oracle.ntt(V)
    for i in (0, 1, 1000, n - 1):
        assert (ev[i] == oracle.poly_eval(V, oracle.fr_pow_u64(d["group_gen"], i))).all()


def test_reference_window_heuristic(pymodel):
    # c and W at the BASELINE sizes
    assert pymodel.ref_window(10) == 3
    assert pymodel.ref_window(1 << 20) == 15 and pymodel.ref_window((1 << 21) - 1) == 16
    assert pymodel.ref_window((1 << 20) + 1) == 16 and pymodel.ref_window(1 << 22) == 17
    assert pymodel.ark_log2(16) == 4 and pymodel.ark_log2(17) == 5 and pymodel.ark_log2(1) == 0
