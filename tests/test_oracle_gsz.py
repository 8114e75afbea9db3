"""Oracle restatement of GSZ20 (mpc-algebra/src/share/gsz20/mod.rs) on the Groth16 path: with the reference's stubbed
preprocessing every share is the plaintext, so the revealed proof must equal the single-prover proof for the same r, s;
the product checks (hadamard_check -> ip_check, field and group) must pass; and beyond the stubs, real Shamir shares
over the mixed-radix share domain must open to the secret and fail the degree check when the degree is too high."""
import random

import numpy as np
import pytest


@pytest.mark.parametrize("parties", [3, 4, 8])  # mpc-algebra/test.zsh runs 3 and 4; BASELINE config 4 is 8 (t = 3)
def test_gsz_proof_equals_plain_and_checks_pass(oracle, pymodel, parties):
    rnd = random.Random(parties)
    n_sq = 13
    toxic = [rnd.randrange(1, pymodel.R_MOD) for _ in range(7)]
    pk = oracle.groth16_setup(n_sq, oracle.fr_from_ints(toxic), threads=4)
    chain = oracle.squaring_chain(oracle.fr_from_ints([rnd.randrange(pymodel.R_MOD)])[0], n_sq)
    one = oracle.fr_from_ints([1])  # the reference's rand() stub: r = s = 1
    res = oracle.groth16_prove_gsz(parties, n_sq, chain, one[0], one[0], pk, threads=2)
    assert res["ok"]
    plain = oracle.groth16_prove(oracle.SCHEME_PLAIN, n_sq, [chain], one, one, pk)
    assert (res["proof"] == plain["proof"]).all() and (res["h"] == plain["h"][0]).all()
    x, y, z = oracle.fr_to_ints(res["field_check"])
    assert x * y % pymodel.R_MOD == z
    gx = oracle.fr_to_ints(res["group_check_x"][None, :])[0]
    gy, gz = oracle.G1.affine_to_ints(res["group_check_yz"], res["group_check_inf"])
    assert pymodel.g1_mul(gy, gx) == gz
    assert res["king_computes"] > 2 * 4 and res["opens"] >= 6
    # other randomness values (a non-stub rand source) still give the plain proof
    r, s = oracle.fr_from_ints([rnd.randrange(pymodel.R_MOD)]), oracle.fr_from_ints([rnd.randrange(pymodel.R_MOD)])
    res = oracle.groth16_prove_gsz(parties, n_sq, chain, r[0], s[0], pk)
    plain = oracle.groth16_prove(oracle.SCHEME_PLAIN, n_sq, [chain], r, s, pk)
    assert res["ok"] and (res["proof"] == plain["proof"]).all()


@pytest.mark.parametrize("parties", [3, 4, 6, 8])
def test_shamir_open_and_degree_check(oracle, pymodel, parties):
    rnd = random.Random(100 + parties)
    t = (parties - 1) // 2
    coeffs = [rnd.randrange(pymodel.R_MOD) for _ in range(t + 1)]
    shares = oracle.gsz_share(parties, oracle.fr_from_ints(coeffs))
    # party j holds p(w^j), w = root of unity of order n (large-subgroup branch)
    w = pymodel.fr_root_of_unity(parties)
    exp = [sum(c * pow(w, j * k, pymodel.R_MOD) for k, c in enumerate(coeffs)) % pymodel.R_MOD for j in range(parties)]
    assert oracle.fr_to_ints(shares) == exp
    val, ok = oracle.gsz_open(shares, t)
    assert ok and oracle.fr_to_ints(val[None, :])[0] == coeffs[0]
    # product of two degree-t sharings has degree 2t: opens at 2t, fails the degree-t check (n > 2t + ... )
    coeffs2 = [rnd.randrange(pymodel.R_MOD) for _ in range(t + 1)]
    prod = oracle.fr_mul(shares, oracle.gsz_share(parties, oracle.fr_from_ints(coeffs2)))
    val, ok = oracle.gsz_open(prod, 2 * t)
    assert ok and oracle.fr_to_ints(val[None, :])[0] == coeffs[0] * coeffs2[0] % pymodel.R_MOD
    if t >= 1:
        _, ok = oracle.gsz_open(prod, t)
        assert not ok
