"""GPU parity for CRS generation on the device (czk_fixed_base_msm, czk_groth16_setup, czk_groth16_setup_r1cs) against the
oracle's restatement of groth16/src/generator.rs:34-221 and FixedBaseMSM (algebra/ec/src/msm/fixed_base.rs:12-96): the same
toxic waste gives bit-identical query points, infinity flags and verifying key; and the whole loop the reference's `proof`
binary runs - setup, prove, verify (mpc-snarks/src/proof.rs:113-141) - closes inside the library."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

QUERIES = (("a_query", "a_inf", 0), ("b_g1_query", "b1_inf", 1), ("b_g2_query", "b2_inf", 2), ("h_query", "h_inf", 3), ("l_query", "l_inf", 4))


def _compare_key(dpk, pk):
    for name, inf_name, which in QUERIES:
        xy, inf = dpk.query(which).numpy()
        assert (inf == pk[inf_name]).all(), name
        keep = pk[inf_name] == 0
        assert (xy[keep] == pk[name][keep]).all(), name


@pytest.mark.parametrize("g", [1, 2])
def test_fixed_base_msm_matches_scalar_mul(ctx, czk, oracle, pymodel, g):
    G = oracle.G1 if g == 1 else oracle.G2
    gen = oracle.generators()[g - 1]
    base = G.scalar_mul(gen, oracle.fr_from_ints([0xabcdef12345])[0])[0]
    vals = [0, 1, 2, 255, 256, pymodel.R_MOD - 1, 1 << 252] + [random.Random(g).randrange(pymodel.R_MOD) for _ in range(60)]
    sc = oracle.fr_from_ints(vals)
    b = czk.fixed_base_msm(ctx, g, base, ctx.vec_from(sc))
    xy, inf = b.numpy()
    assert inf[0] == 1 and not inf[1:].any()
    for i in range(1, len(vals)):
        exp, einf = G.scalar_mul(base, sc[i])
        assert not einf and (xy[i] == exp).all(), i


@pytest.mark.parametrize("n_sq", [10, 1 << 10])
def test_device_setup_matches_oracle_and_loop_closes(ctx, czk, oracle, pymodel, n_sq):
    ctx.net_init(0, 1, None)
    rnd = random.Random(n_sq)
    toxic = oracle.fr_from_ints([rnd.randrange(1, pymodel.R_MOD) for _ in range(7)])
    pk = oracle.groth16_setup(n_sq, toxic, threads=oracle.cpu_threads())
    dpk = czk.groth16_setup(ctx, n_sq, toxic)
    assert dpk.domain_size == pk["D"]
    _compare_key(dpk, pk)
    vk = czk.pk_verifying_key(dpk)
    assert (vk["vk_g1"] == pk["vk_g1"]).all() and (vk["vk_g2"] == pk["vk_g2"]).all() and (vk["gamma_abc_g1"] == pk["gamma_abc_g1"]).all()
    # setup -> prove (SPDZ shares, one party) -> verify, all through the library
    chain = czk.squaring_chain(oracle.fr_from_ints([rnd.randrange(pymodel.R_MOD)])[0], n_sq)
    r, s = oracle.random_fr_mont(1, 1)[0], oracle.random_fr_mont(2, 1)[0]
    got = czk.groth16_prove(ctx, czk.SCHEME_SPDZ, dpk, chain, r, s)
    assert czk.groth16_verify(vk, chain[n_sq:n_sq + 1], got["proof"], got["proof_inf"])
    assert not czk.groth16_verify(vk, chain[0:1], got["proof"], got["proof_inf"])
    dpk.free()


def test_device_setup_any_circuit(ctx, czk, oracle, pymodel):
    ctx.net_init(0, 1, None)
    rnd = random.Random(5)
    cs, z = oracle.random_r1cs(seed=12, n_inst=4, n_free=10, n_cons=500, modulus=pymodel.R_MOD)
    toxic = oracle.fr_from_ints([rnd.randrange(1, pymodel.R_MOD) for _ in range(7)])
    pk = oracle.groth16_setup_r1cs(cs, toxic, threads=oracle.cpu_threads())
    dpk = czk.groth16_setup_r1cs(ctx, cs, toxic)
    _compare_key(dpk, pk)
    vk = czk.pk_verifying_key(dpk)
    assert (vk["gamma_abc_g1"] == pk["gamma_abc_g1"]).all()
    dcs = czk.R1cs(ctx, cs)
    full = oracle.r1cs_full_shares(z, 1, seed=1, scheme=oracle.SCHEME_ADDITIVE)
    got = czk.groth16_prove_r1cs(ctx, czk.SCHEME_ADDITIVE, dpk, dcs, full[0], oracle.random_fr_mont(3, 1)[0], oracle.random_fr_mont(4, 1)[0])
    assert czk.groth16_verify(vk, oracle.fr_from_ints(z[1:4]), got["proof"], got["proof_inf"])
    dcs.free()
    dpk.free()


def test_real_key_proof_verifies_at_2_16(ctx, czk, oracle):
    """A real CRS at 2^16 constraints generated on the device in well under a second; the SPDZ proof verifies."""
    ctx.net_init(0, 1, None)
    n_sq = 1 << 16
    toxic = oracle.random_fr_mont(99, 7)
    dpk = czk.groth16_setup(ctx, n_sq, toxic)
    chain = czk.squaring_chain(oracle.random_fr_mont(98, 1)[0], n_sq)
    got = czk.groth16_prove(ctx, czk.SCHEME_SPDZ, dpk, chain, oracle.random_fr_mont(97, 1)[0], oracle.random_fr_mont(96, 1)[0])
    assert czk.groth16_verify(czk.pk_verifying_key(dpk), chain[n_sq:n_sq + 1], got["proof"], got["proof_inf"])
    dpk.free()
