"""Pin the oracle's Groth16 restatement (oracle/czk_oracle_groth16.inc): the proofs it produces satisfy the
Groth16 verification equation.  The reference checks every benchmark proof with verify_proof
(mpc-snarks/src/proof.rs:141, groth16/src/test.rs:78-108); a pairing is out of scope here, so the same
equation  e(A,B) = e(alpha,beta) e(sum x_i gamma_abc_i, gamma) e(C, delta)  is checked in the exponent,
which the known toxic waste allows: A = a*g1, B = b*g2, C = c*g1 with
    a*b == alpha*beta + (sum_i x_i abc_i)*gamma... (all mod r),
and the discrete logs a, b, c are recomputed independently with Python big ints from the QAP.
Also: MPC proofs (additive, SPDZ; 2 and 3 parties) reveal to exactly the single-prover proof for r = sum r_p,
s = sum s_p, and a wrong public input is rejected (groth16/src/test.rs:158-171)."""
import random

import numpy as np
import pytest


def qap_at_tau(pymodel, n_sq, tau):
    """a_i(tau), b_i(tau), c_i(tau) for variables [one, out, w_0..w_{n-1}] (groth16/src/r1cs_to_qap.rs:51-92)."""
    m = pymodel
    d = m.Domain(n_sq + 2)
    D, R = d.size, m.R_MOD
    zt = d.vanishing_at(tau)
    u = [zt * pow(d.group_gen, i, R) * pow(D * (tau - pow(d.group_gen, i, R)), -1, R) % R for i in range(D)]
    nv = n_sq + 2
    a, b, c = [0] * nv, [0] * nv, [0] * nv
    a[0], a[1] = u[n_sq], u[n_sq + 1]
    for i in range(n_sq):
        a[2 + i] = (a[2 + i] + u[i]) % R
        b[2 + i] = (b[2 + i] + u[i]) % R
        ci = 2 + i + 1 if i + 1 < n_sq else 1
        c[ci] = (c[ci] + u[i]) % R
    return a, b, c, zt, D


def check_in_exponent(pymodel, oracle, n_sq, toxic, chain, r, s, h_coeffs, proof, proof_inf, public_out=None):
    m, R = pymodel, pymodel.R_MOD
    alpha, beta, gamma, delta, tau, s1, s2 = toxic
    a, b, c, zt, D = qap_at_tau(m, n_sq, tau)
    z = [1, chain[n_sq]] + chain[:n_sq]
    A = (alpha + sum(x * y for x, y in zip(z, a)) + r * delta) % R
    B = (beta + sum(x * y for x, y in zip(z, b)) + s * delta) % R
    dinv = pow(delta, -1, R)
    l = [(beta * a[i] + alpha * b[i] + c[i]) * dinv % R for i in range(n_sq + 2)]
    h_tau = sum(hc * pow(tau, i, R) for i, hc in enumerate(h_coeffs[:D - 1])) % R
    Cc = (sum(z[i] * l[i] for i in range(2, n_sq + 2)) + h_tau * zt * dinv + s * A + r * B - r * s * delta) % R
    # the prover's group elements are exactly these multiples of the CRS generators
    g1 = m.g1_mul(m.G1_GEN, s1)
    g2 = m.g2_mul(m.G2_GEN, s2)
    pa = oracle.G1.affine_to_ints(proof[:12])[0]
    pb = oracle.G2.affine_to_ints(proof[12:36])[0]
    pc = oracle.G1.affine_to_ints(proof[36:48])[0]
    assert not proof_inf.any()
    assert pa == m.g1_mul(g1, A) and pb == m.g2_mul(g2, B) and pc == m.g1_mul(g1, Cc)
    # Groth16 verification equation in the exponent
    ginv = pow(gamma, -1, R)
    abc = [(beta * a[i] + alpha * b[i] + c[i]) * ginv % R for i in range(2)]
    x = [1, chain[n_sq] if public_out is None else public_out]
    lhs = A * B % R
    rhs = (alpha * beta + sum(xi * ai for xi, ai in zip(x, abc)) * gamma + Cc * delta) % R
    return lhs == rhs


@pytest.fixture(scope="module")
def setup10(oracle, pymodel):
    rnd = random.Random(42)
    n_sq = 10  # BASELINE config 1, literal: `bench.zsh groth16 spdz 10 2`
    toxic = [rnd.randrange(1, pymodel.R_MOD) for _ in range(7)]
    pk = oracle.groth16_setup(n_sq, oracle.fr_from_ints(toxic), threads=4)
    start = rnd.randrange(pymodel.R_MOD)
    chain_m = oracle.squaring_chain(oracle.fr_from_ints([start])[0], n_sq)
    chain = oracle.fr_to_ints(chain_m)
    assert chain == [pow(start, 1 << i, pymodel.R_MOD) for i in range(n_sq + 1)]
    return dict(n_sq=n_sq, toxic=toxic, pk=pk, chain=chain, chain_m=chain_m, rnd=rnd)


def test_crs_shape_and_infinity_entries(setup10, oracle):
    pk = setup10["pk"]
    assert pk["D"] == 16 and pk["h_query"].shape[0] == 15
    # b has no term for `one` and `out`: those B-query entries are the point at infinity
    assert list(pk["b1_inf"][:2]) == [1, 1] and not pk["b1_inf"][2:].any()
    assert list(pk["b2_inf"][:2]) == [1, 1]
    assert not pk["a_inf"].any() and not pk["h_inf"].any()


def test_plain_proof_verifies_in_exponent(setup10, oracle, pymodel):
    S = setup10
    rnd = random.Random(1)
    r, s = rnd.randrange(pymodel.R_MOD), rnd.randrange(pymodel.R_MOD)
    res = oracle.groth16_prove(oracle.SCHEME_PLAIN, S["n_sq"], [S["chain_m"]], oracle.fr_from_ints([r]), oracle.fr_from_ints([s]), S["pk"])
    assert res["ok"]
    h = oracle.fr_to_ints(res["h"][0])
    assert h[-1] == 0  # deg h <= D - 2
    assert check_in_exponent(pymodel, oracle, S["n_sq"], S["toxic"], S["chain"], r, s, h, res["proof"], res["proof_inf"])
    # negative: wrong public input fails the equation (groth16/src/test.rs:158-171)
    assert not check_in_exponent(pymodel, oracle, S["n_sq"], S["toxic"], S["chain"], r, s, h, res["proof"], res["proof_inf"],
                                 public_out=(S["chain"][-1] + 1) % pymodel.R_MOD)


@pytest.mark.parametrize("scheme_name,parties", [("additive", 2), ("spdz", 2), ("spdz", 3), ("additive", 4)])
def test_mpc_proof_reveals_to_plain_proof(setup10, oracle, pymodel, scheme_name, parties):
    S = setup10
    scheme = oracle.SCHEME_ADDITIVE if scheme_name == "additive" else oracle.SCHEME_SPDZ
    shares = oracle.king_share_batch(S["chain_m"], parties, seed=7)
    # MpcField::rand: every party draws the same value from the same seeded rng (wire/macros.rs:129-131)
    rho, sigma = 123456789, 987654321
    r_sh = oracle.fr_from_ints([rho] * parties)
    s_sh = oracle.fr_from_ints([sigma] * parties)
    res = oracle.groth16_prove(scheme, S["n_sq"], shares, r_sh, s_sh, S["pk"], threads=2)
    assert res["ok"], "MAC check failed"
    r, s = rho * parties % pymodel.R_MOD, sigma * parties % pymodel.R_MOD
    plain = oracle.groth16_prove(oracle.SCHEME_PLAIN, S["n_sq"], [S["chain_m"]], oracle.fr_from_ints([r]), oracle.fr_from_ints([s]), S["pk"])
    assert (res["proof"] == plain["proof"]).all() and (res["proof_inf"] == plain["proof_inf"]).all()
    # shares of h sum to the plain h (the Beaver product is exact)
    hsum = res["h"][0]
    for p in range(1, parties):
        hsum = oracle.fr_add(hsum, res["h"][p])
    assert (hsum == plain["h"][0]).all()
    h = oracle.fr_to_ints(plain["h"][0])
    assert check_in_exponent(pymodel, oracle, S["n_sq"], S["toxic"], S["chain"], r, s, h, res["proof"], res["proof_inf"])


def test_witness_map_only_matches_prover_h(setup10, oracle):
    S = setup10
    shares = oracle.king_share_batch(S["chain_m"], 2, seed=9)
    h, ok = oracle.groth16_witness_map(oracle.SCHEME_SPDZ, S["n_sq"], shares, threads=2)
    res = oracle.groth16_prove(oracle.SCHEME_SPDZ, S["n_sq"], shares, oracle.fr_from_ints([1, 1]), oracle.fr_from_ints([2, 2]), S["pk"])
    assert ok and (h == res["h"]).all()


def test_larger_circuit_2_8(oracle, pymodel):
    rnd = random.Random(5)
    n_sq = 1 << 8
    toxic = [rnd.randrange(1, pymodel.R_MOD) for _ in range(7)]
    pk = oracle.groth16_setup(n_sq, oracle.fr_from_ints(toxic), threads=oracle.cpu_threads())
    assert pk["D"] == 512
    start = rnd.randrange(pymodel.R_MOD)
    chain_m = oracle.squaring_chain(oracle.fr_from_ints([start])[0], n_sq)
    r, s = rnd.randrange(pymodel.R_MOD), rnd.randrange(pymodel.R_MOD)
    res = oracle.groth16_prove(oracle.SCHEME_PLAIN, n_sq, [chain_m], oracle.fr_from_ints([r]), oracle.fr_from_ints([s]), pk, threads=4)
    h = oracle.fr_to_ints(res["h"][0])
    assert check_in_exponent(pymodel, oracle, n_sq, toxic, oracle.fr_to_ints(chain_m), r, s, h, res["proof"], res["proof_inf"])


# ------------------------------------------------------------------ any R1CS, not only the benchmark's squaring chain
def qap_at_tau_r1cs(pymodel, cs, oracle, tau):
    m, R = pymodel, pymodel.R_MOD
    d = m.Domain(cs["ncons"] + cs["ninst"])
    D = d.size
    zt = d.vanishing_at(tau)
    u = [zt * pow(d.group_gen, i, R) * pow(D * (tau - pow(d.group_gen, i, R)), -1, R) % R for i in range(D)]
    nv = cs["ninst"] + cs["nwit"]
    out = {}
    for name in ("a", "b", "c"):
        rp, col, cf = cs[name]
        cfi = oracle.fr_to_ints(cf)
        v = [0] * nv
        for i in range(cs["ncons"]):
            for k in range(int(rp[i]), int(rp[i + 1])):
                v[int(col[k])] = (v[int(col[k])] + u[i] * cfi[k]) % R
        out[name] = v
    for i in range(cs["ninst"]):
        out["a"][i] = (out["a"][i] + u[cs["ncons"] + i]) % R
    return out["a"], out["b"], out["c"], zt, D


@pytest.mark.parametrize("scheme_name,parties", [("plain", 1), ("additive", 2), ("spdz", 3)])
def test_generic_r1cs_proof_verifies_in_exponent(oracle, pymodel, scheme_name, parties):
    """groth16/src/test.rs:78-108 on a random circuit (3 instance variables, rows of 1-3 terms, arbitrary coefficients):
    the generic prover core is the one the squaring entry points use, so this pins it beyond the benchmark circuit."""
    m, R = pymodel, pymodel.R_MOD
    rnd = random.Random(99)
    cs, z = oracle.random_r1cs(seed=5, n_inst=3, n_free=4, n_cons=21, modulus=R)
    toxic = [rnd.randrange(1, R) for _ in range(7)]
    alpha, beta, gamma, delta, tau, s1, s2 = toxic
    pk = oracle.groth16_setup_r1cs(cs, oracle.fr_from_ints(toxic), threads=2)
    scheme = {"plain": oracle.SCHEME_PLAIN, "additive": oracle.SCHEME_ADDITIVE, "spdz": oracle.SCHEME_SPDZ}[scheme_name]
    full = oracle.r1cs_full_shares(z, parties, seed=3, scheme=scheme)
    rho, sigma = rnd.randrange(R), rnd.randrange(R)
    res = oracle.groth16_prove_r1cs(scheme, cs, full, oracle.fr_from_ints([rho] * parties), oracle.fr_from_ints([sigma] * parties), pk)
    assert res["ok"]
    r, s = rho * parties % R, sigma * parties % R
    hsum = res["h"][0]
    for p in range(1, parties):
        hsum = oracle.fr_add(hsum, res["h"][p])
    h = oracle.fr_to_ints(hsum)
    a, b, c, zt, D = qap_at_tau_r1cs(m, cs, oracle, tau)
    nv, ninst = len(z), cs["ninst"]
    A = (alpha + sum(x * y for x, y in zip(z, a)) + r * delta) % R
    B = (beta + sum(x * y for x, y in zip(z, b)) + s * delta) % R
    dinv, ginv = pow(delta, -1, R), pow(gamma, -1, R)
    l = [(beta * a[i] + alpha * b[i] + c[i]) * dinv % R for i in range(nv)]
    h_tau = sum(hc * pow(tau, i, R) for i, hc in enumerate(h[:D - 1])) % R
    Cc = (sum(z[i] * l[i] for i in range(ninst, nv)) + h_tau * zt * dinv + s * A + r * B - r * s * delta) % R
    g1, g2 = m.g1_mul(m.G1_GEN, s1), m.g2_mul(m.G2_GEN, s2)
    assert not res["proof_inf"].any()
    assert oracle.G1.affine_to_ints(res["proof"][:12])[0] == m.g1_mul(g1, A)
    assert oracle.G2.affine_to_ints(res["proof"][12:36])[0] == m.g2_mul(g2, B)
    assert oracle.G1.affine_to_ints(res["proof"][36:48])[0] == m.g1_mul(g1, Cc)
    abc = [(beta * a[i] + alpha * b[i] + c[i]) * ginv % R for i in range(ninst)]
    assert A * B % R == (alpha * beta + sum(z[i] * abc[i] for i in range(ninst)) * gamma + Cc * delta) % R
    # a wrong public input breaks the equation
    bad = list(z[:ninst])
    bad[1] = (bad[1] + 1) % R
    assert A * B % R != (alpha * beta + sum(bad[i] * abc[i] for i in range(ninst)) * gamma + Cc * delta) % R
