"""CPU tier: the product's host-side BLS12-377 pairing-product check and Groth16 verifier (csrc/pairing.cu) against
(a) the oracle's Python big-int model of the same ate pairing (oracle/pymodel.py), (b) bilinearity, and (c) oracle proofs
whose validity is established independently in the exponent (tests/test_oracle_groth16.py) - the reference's own
acceptance test is `verify_proof` after every proof (mpc-snarks/src/proof.rs:141, groth16/src/test.rs:78-108,158-171)."""
import random

import numpy as np
import pytest


def test_pairing_product_bilinearity_and_model(czk, oracle, pymodel):
    m = pymodel
    rnd = random.Random(3)
    a, b = rnd.randrange(1, m.R_MOD), rnd.randrange(1, m.R_MOD)
    P, Q = m.G1_GEN, m.G2_GEN
    aP, bQ, abP = m.g1_mul(P, a), m.g2_mul(Q, b), m.g1_mul(P, a * b % m.R_MOD)
    g1 = lambda pts: oracle.G1.affine_from_ints(pts)[0]
    g2 = lambda pts: oracle.G2.affine_from_ints(pts)[0]
    # e(aP, bQ) * e(-abP, Q) == 1
    assert czk.pairing_product_is_one(g1([aP, m.g1_neg(abP)]), g2([bQ, Q]))
    assert m.pairing_product_is_one([(aP, bQ), (m.g1_neg(abP), Q)])
    # ... and not when one factor is off by one
    off = m.g1_neg(m.g1_mul(P, (a * b + 1) % m.R_MOD))
    assert not czk.pairing_product_is_one(g1([aP, off]), g2([bQ, Q]))
    assert not m.pairing_product_is_one([(aP, bQ), (off, Q)])
    # non-degeneracy: a single pairing of the generators is not 1; pairs with infinity contribute 1
    assert not czk.pairing_product_is_one(g1([P]), g2([Q]))
    assert czk.pairing_product_is_one(g1([P]), g2([Q]), g1_inf=np.ones(1, np.uint8))


@pytest.mark.parametrize("circuit", ["squaring", "random"])
def test_groth16_verify_accepts_oracle_proofs_and_rejects_tampering(czk, oracle, pymodel, circuit):
    m, R = pymodel, pymodel.R_MOD
    rnd = random.Random(11)
    toxic = oracle.fr_from_ints([rnd.randrange(1, R) for _ in range(7)])
    r, s = oracle.fr_from_ints([rnd.randrange(R)]), oracle.fr_from_ints([rnd.randrange(R)])
    if circuit == "squaring":
        n_sq = 10
        pk = oracle.groth16_setup(n_sq, toxic, threads=2)
        chain = oracle.squaring_chain(oracle.fr_from_ints([rnd.randrange(R)])[0], n_sq)
        res = oracle.groth16_prove(oracle.SCHEME_PLAIN, n_sq, [chain], r, s, pk)
        public = chain[n_sq:n_sq + 1]
    else:
        cs, z = oracle.random_r1cs(seed=4, n_inst=3, n_free=5, n_cons=30, modulus=R)
        pk = oracle.groth16_setup_r1cs(cs, toxic, threads=2)
        res = oracle.groth16_prove_r1cs(oracle.SCHEME_PLAIN, cs, oracle.r1cs_full_shares(z, 1, 0, oracle.SCHEME_PLAIN), r, s, pk)
        public = oracle.fr_from_ints(z[1:3])
    assert res["ok"]
    assert czk.groth16_verify(pk, public, res["proof"], res["proof_inf"])
    # the Python model agrees
    G1, G2 = oracle.G1, oracle.G2
    vk = dict(alpha_g1=G1.affine_to_ints(pk["vk_g1"][0:1])[0], beta_g2=G2.affine_to_ints(pk["vk_g2"][0:1])[0],
              gamma_g2=G2.affine_to_ints(pk["vk_g2"][1:2])[0], delta_g2=G2.affine_to_ints(pk["vk_g2"][2:3])[0],
              gamma_abc_g1=G1.affine_to_ints(pk["gamma_abc_g1"]))
    proof = (G1.affine_to_ints(res["proof"][None, :12])[0], G2.affine_to_ints(res["proof"][None, 12:36])[0],
             G1.affine_to_ints(res["proof"][None, 36:48])[0])
    assert m.groth16_verify(vk, proof, oracle.fr_to_ints(public))
    # wrong public input (groth16/src/test.rs:158-171), and a proof element swapped for another group element
    bad = oracle.fr_add(public, oracle.fr_from_ints([1] * public.shape[0]))
    assert not czk.groth16_verify(pk, bad, res["proof"], res["proof_inf"])
    tampered = res["proof"].copy()
    tampered[36:48] = pk["vk_g1"][1]
    assert not czk.groth16_verify(pk, public, tampered, res["proof_inf"])
    # the proof survives the wire format
    back, inf = czk.groth16_proof_deserialize(czk.groth16_proof_serialize(res["proof"], res["proof_inf"]))
    assert czk.groth16_verify(pk, public, back, inf)


def test_verify_rejects_points_off_the_curve_or_subgroup(czk, oracle, pymodel):
    """czk_groth16_verify on raw limbs checks what the reference's typed, deserialised points guarantee: a proof element
    that is not a point of the prime-order subgroup is an argument error (CZK_ERR_ARG), not a verdict."""
    import random

    rnd = random.Random(9)
    n_sq = 4
    toxic = [rnd.randrange(1, pymodel.R_MOD) for _ in range(7)]
    pk = oracle.groth16_setup(n_sq, oracle.fr_from_ints(toxic))
    chain = oracle.squaring_chain(oracle.fr_from_ints([rnd.randrange(pymodel.R_MOD)])[0], n_sq)
    r, s = oracle.fr_from_ints([5]), oracle.fr_from_ints([7])
    pf = oracle.groth16_prove(oracle.SCHEME_PLAIN, n_sq, [chain], r, s, pk)
    proof, inf = pf["proof"].copy(), pf["proof_inf"].copy()
    assert czk.groth16_verify(pk, chain[n_sq:n_sq + 1], proof, inf)
    bad = proof.copy()
    bad[6] ^= np.uint64(1)  # A.y: no longer on the curve
    with pytest.raises(czk.CzkError) as e:
        czk.groth16_verify(pk, chain[n_sq:n_sq + 1], bad, inf)
    assert e.value.code == 2
