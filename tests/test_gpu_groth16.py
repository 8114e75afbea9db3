"""GPU parity for the Groth16 prover loop (czk_groth16_prove / czk_groth16_witness_map) against the oracle's
restatement of mpc-snarks/src/groth/{prover.rs,r1cs_to_qap.rs}: same CRS, same witness shares, same r and s
=> bit-identical h shares and bit-identical affine proof elements.  One party here (1 GPU); the multi-party
run is tests/mp_groth16_check.py (torchrun, one rank per GPU)."""
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


@pytest.mark.parametrize("n_sq", [10, 1 << 10])  # BASELINE config 1: literal `... spdz 10 2` and the described 2^10
@pytest.mark.parametrize("scheme_name", ["plain", "additive", "spdz"])
def test_prove_matches_oracle_single_party(ctx, czk, oracle, pymodel, n_sq, scheme_name):
    ctx.net_init(0, 1, None)
    scheme = {"plain": czk.SCHEME_PLAIN, "additive": czk.SCHEME_ADDITIVE, "spdz": czk.SCHEME_SPDZ}[scheme_name]
    pk, chain, r, s = _setup(oracle, pymodel, n_sq, seed=n_sq)
    exp = oracle.groth16_prove(scheme, n_sq, [chain], r, s, pk, threads=oracle.cpu_threads())
    assert exp["ok"]
    dpk = czk.ProvingKey.upload(ctx, pk)
    assert dpk.domain_size == pk["D"]
    h = czk.groth16_witness_map(ctx, scheme, n_sq, chain)
    assert (h == exp["h"][0]).all()
    got = czk.groth16_prove(ctx, scheme, dpk, chain, r[0], s[0])
    assert (got["proof_inf"] == exp["proof_inf"]).all() and (got["proof"] == exp["proof"]).all()
    assert (got["proof_sh"] == exp["proof_sh"][0]).all() and (got["proof_sh_inf"] == exp["proof_sh_inf"][0]).all()
    # the reference's acceptance test (proof.rs:141 verify_proof): the device proof verifies under the pairing check,
    # survives Proof::serialize, and is rejected for another public input (groth16/src/test.rs:158-171)
    public = chain[n_sq:n_sq + 1]
    wire = czk.groth16_proof_serialize(got["proof"], got["proof_inf"])
    assert len(wire) == 192
    back, binf = czk.groth16_proof_deserialize(wire)
    assert czk.groth16_verify(pk, public, back, binf)
    assert not czk.groth16_verify(pk, oracle.fr_add(public, oracle.fr_from_ints([1])), got["proof"], got["proof_inf"])
    dpk.free()


def test_witness_map_at_baseline_domain_2_21(ctx, czk, oracle):
    """Config C2 domain: N = 2^20 squarings => D = 2^21; the full NTT + product pipeline, bit-exact."""
    ctx.net_init(0, 1, None)
    n_sq = 1 << 20
    chain = oracle.squaring_chain(oracle.random_fr_mont(77, 1)[0], n_sq)
    exp, ok = oracle.groth16_witness_map(oracle.SCHEME_SPDZ, n_sq, [chain], threads=oracle.cpu_threads())
    assert ok
    got = czk.groth16_witness_map(ctx, czk.SCHEME_SPDZ, n_sq, chain)
    assert got.shape == (1 << 21, 4) and (got == exp[0]).all()


def test_prove_synthetic_key_2_14_matches_oracle(ctx, czk, oracle):
    """The benchmark's synthetic (device-generated) key, downloaded and fed to the oracle prover."""
    ctx.net_init(0, 1, None)
    n_sq = 1 << 14
    dpk = czk.ProvingKey.synthetic(ctx, n_sq, seed=3)
    pk = dpk.to_host()
    chain = oracle.squaring_chain(oracle.random_fr_mont(5, 1)[0], n_sq)
    r, s = oracle.random_fr_mont(6, 1), oracle.random_fr_mont(7, 1)
    exp = oracle.groth16_prove(oracle.SCHEME_SPDZ, n_sq, [chain], r, s, pk, threads=oracle.cpu_threads(), want_h=False)
    got = czk.groth16_prove(ctx, czk.SCHEME_SPDZ, dpk, chain, r[0], s[0])
    assert (got["proof"] == exp["proof"]).all() and (got["proof_inf"] == exp["proof_inf"]).all()
    dpk.free()


@pytest.mark.parametrize("scheme_name", ["plain", "additive", "spdz", "gsz"])
@pytest.mark.parametrize("shape", [(3, 4, 21), (2, 50, 3000)])  # (instance variables, free witnesses, constraints)
def test_prove_any_r1cs_matches_oracle(ctx, czk, oracle, pymodel, scheme_name, shape):
    """czk_groth16_prove_r1cs: constraint matrices in CSR form, sparse evaluate_constraint on the device
    (mpc-snarks/src/groth/r1cs_to_qap.rs:12-41,70-83), then the same prover loop.  Random satisfiable circuits with rows
    of 1-3 terms and arbitrary coefficients; the oracle's generic core is pinned by the exponent check in
    tests/test_oracle_groth16.py.  One party here; tests/mp_groth16_check.py covers 2+ parties."""
    ctx.net_init(0, 1, None)
    ninst, nfree, ncons = shape
    rnd = random.Random(ncons)
    cs, z = oracle.random_r1cs(seed=ncons, n_inst=ninst, n_free=nfree, n_cons=ncons, modulus=pymodel.R_MOD)
    toxic = [rnd.randrange(1, pymodel.R_MOD) for _ in range(7)]
    pk = oracle.groth16_setup_r1cs(cs, oracle.fr_from_ints(toxic), threads=oracle.cpu_threads())
    r = oracle.fr_from_ints([rnd.randrange(pymodel.R_MOD)])
    s = oracle.fr_from_ints([rnd.randrange(pymodel.R_MOD)])
    dcs = czk.R1cs(ctx, cs)
    dpk = czk.groth16_pk_upload_r1cs(ctx, pk)
    if scheme_name == "gsz":
        # GSZ: every party holds the plaintext (gsz20/mod.rs:202-212); the revealed proof is the single-prover proof
        full = oracle.r1cs_full_shares(z, 1, seed=1, scheme=oracle.SCHEME_PLAIN)
        exp = oracle.groth16_prove_r1cs(oracle.SCHEME_PLAIN, cs, full, r, s, pk, threads=oracle.cpu_threads())
        got = czk.groth16_prove_r1cs(ctx, czk.SCHEME_GSZ, dpk, dcs, full[0], r[0], s[0])
    else:
        oscheme = {"plain": oracle.SCHEME_PLAIN, "additive": oracle.SCHEME_ADDITIVE, "spdz": oracle.SCHEME_SPDZ}[scheme_name]
        scheme = {"plain": czk.SCHEME_PLAIN, "additive": czk.SCHEME_ADDITIVE, "spdz": czk.SCHEME_SPDZ}[scheme_name]
        full = oracle.r1cs_full_shares(z, 1, seed=1, scheme=oscheme)
        exp = oracle.groth16_prove_r1cs(oscheme, cs, full, r, s, pk, threads=oracle.cpu_threads())
        got = czk.groth16_prove_r1cs(ctx, scheme, dpk, dcs, full[0], r[0], s[0])
        assert (got["proof_sh"] == exp["proof_sh"][0]).all()
    assert exp["ok"]
    assert (got["proof"] == exp["proof"]).all() and (got["proof_inf"] == exp["proof_inf"]).all()
    assert czk.groth16_verify(pk, oracle.fr_from_ints(z[1:ninst]), got["proof"], got["proof_inf"])
    # the squaring entry points refuse a key of another circuit
    with pytest.raises(czk.CzkError):
        czk.groth16_prove(ctx, czk.SCHEME_PLAIN, dpk, np.zeros((1, 4), np.uint64), r[0], s[0])
    dcs.free()
    dpk.free()
