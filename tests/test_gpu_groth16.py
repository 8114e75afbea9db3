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
