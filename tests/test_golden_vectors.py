"""Frozen known-answer vectors (tests/golden/vectors.json, written by tests/golden/make_vectors.py): the oracle must keep
reproducing them (CPU tier), and the device path must hit the same values (GPU tier)."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from helpers import jac_to_affine_ints, make_points

VEC = json.loads((Path(__file__).resolve().parent / "golden" / "vectors.json").read_text())


def digest(arr) -> str:
    return hashlib.sha256(np.ascontiguousarray(arr, np.uint64).tobytes()).hexdigest()


def _pt(hexes, g):
    if hexes is None:
        return None
    v = [int(h, 16) for h in hexes]
    return (v[0], v[1]) if g == "g1" else ((v[0], v[1]), (v[2], v[3]))


def _msm_inputs(oracle, e):
    G = oracle.G1 if e["group"] == "g1" else oracle.G2
    xy = make_points(G, e["n"], seed=e["points_seed"])
    sc = oracle.random_fr_mont(e["scalars_seed"], e["n"])
    inf = np.zeros(e["n"], np.uint8)
    inf[::e["inf_every"]] = 1
    return G, xy, inf, sc


@pytest.mark.parametrize("e", VEC["msm"], ids=lambda e: f"{e['group']}-{e['n']}")
def test_oracle_msm_golden(oracle, e):
    G, xy, inf, sc = _msm_inputs(oracle, e)
    res, isinf = G.msm(xy, inf, sc, threads=2)
    assert (None if isinf else G.affine_to_ints(res)[0]) == _pt(e["result"], e["group"])


@pytest.mark.parametrize("e", VEC["ntt"], ids=lambda e: f"2^{e['log_d']}")
def test_oracle_ntt_golden(oracle, e):
    v = oracle.random_fr_mont(e["seed"], 1 << e["log_d"])
    for name, inv, cos in (("fft", False, False), ("ifft", True, False), ("coset_fft", False, True), ("coset_ifft", True, True)):
        assert digest(oracle.ntt(v, inv, cos)) == e[name], name


def _groth_inputs(oracle, n_sq):
    toxic = oracle.random_fr_mont(31, 7)
    chain = oracle.squaring_chain(oracle.random_fr_mont(32, 1)[0], n_sq)
    return toxic, chain, oracle.random_fr_mont(33, 1), oracle.random_fr_mont(34, 1)


@pytest.mark.parametrize("e", VEC["groth16"], ids=lambda e: f"n={e['n_sq']}")
def test_oracle_groth16_golden(oracle, e):
    toxic, chain, r, s = _groth_inputs(oracle, e["n_sq"])
    pk = oracle.groth16_setup(e["n_sq"], toxic, threads=2)
    assert digest(np.concatenate([pk[k].reshape(-1) for k in ("a_query", "b_g1_query", "b_g2_query", "h_query", "l_query")])) == e["pk_digest"]
    res = oracle.groth16_prove(oracle.SCHEME_PLAIN, e["n_sq"], [chain], r, s, pk, threads=2)
    assert digest(res["proof"]) == e["proof"] and digest(res["h"][0]) == e["h"]


@pytest.mark.gpu
@pytest.mark.parametrize("e", VEC["msm"], ids=lambda e: f"{e['group']}-{e['n']}")
def test_device_msm_golden(ctx, oracle, e):
    G, xy, inf, sc = _msm_inputs(oracle, e)
    fn = ctx.msm_g1 if e["group"] == "g1" else ctx.msm_g2
    assert jac_to_affine_ints(G, fn(xy, inf, sc, True)) == _pt(e["result"], e["group"])


@pytest.mark.gpu
@pytest.mark.parametrize("e", VEC["ntt"], ids=lambda e: f"2^{e['log_d']}")
def test_device_ntt_golden(ctx, oracle, e):
    v = oracle.random_fr_mont(e["seed"], 1 << e["log_d"])
    for name, inv, cos in (("fft", False, False), ("ifft", True, False), ("coset_fft", False, True), ("coset_ifft", True, True)):
        assert digest(ctx._ntt_host(v, inv, cos)) == e[name], name


@pytest.mark.gpu
@pytest.mark.parametrize("e", VEC["groth16"], ids=lambda e: f"n={e['n_sq']}")
def test_device_groth16_golden(ctx, czk, oracle, e):
    """Key generated on the device, witness map and proof on the device: all three digests are the frozen ones."""
    ctx.net_init(0, 1, None)
    toxic, chain, r, s = _groth_inputs(oracle, e["n_sq"])
    dpk = czk.groth16_setup(ctx, e["n_sq"], toxic)
    host = dpk.to_host()
    assert digest(np.concatenate([host[k].reshape(-1) for k in ("a_query", "b_g1_query", "b_g2_query", "h_query", "l_query")])) == e["pk_digest"]
    assert digest(czk.groth16_witness_map(ctx, czk.SCHEME_PLAIN, e["n_sq"], chain)) == e["h"]
    got = czk.groth16_prove(ctx, czk.SCHEME_PLAIN, dpk, chain, r[0], s[0])
    assert digest(got["proof"]) == e["proof"]
    dpk.free()


# ---------------------------------------------------------------------------------------------------------------------
# A second, independent implementation against the SAME frozen values: oracle/pymodel.py (Python big integers, textbook
# affine group law, naive sum_i s_i P_i as in algebra/test-templates/src/msm.rs:6-14, a plain DFT) must reproduce every
# golden MSM result and NTT digest.  The C oracle and this model share no code and no algorithm (Pippenger vs naive,
# in-place radix-2 with derange vs recursive DFT), so agreement on the frozen vectors pins both to the mathematics.
@pytest.mark.parametrize("e", VEC["msm"], ids=lambda e: f"{e['group']}-{e['n']}")
def test_pymodel_msm_golden(oracle, pymodel, e):
    G, xy, inf, sc = _msm_inputs(oracle, e)
    pts = G.affine_to_ints(xy, inf)
    scalars = oracle.fr_to_ints(sc)  # canonical integers
    if e["group"] == "g1":
        acc = None
        for P, s in zip(pts, scalars):
            if P is not None and s:
                acc = pymodel.g1_add(acc, pymodel.g1_mul(P, s))
    else:
        acc = None
        for P, s in zip(pts, scalars):
            if P is not None and s:
                acc = pymodel.g2_add(acc, pymodel.g2_mul(P, s))
    assert acc == _pt(e["result"], e["group"])


@pytest.mark.parametrize("e", VEC["ntt"], ids=lambda e: f"2^{e['log_d']}")
def test_pymodel_ntt_golden(oracle, pymodel, e):
    v = oracle.random_fr_mont(e["seed"], 1 << e["log_d"])
    ints = oracle.fr_to_ints(v)
    for name, inv, cos in (("fft", False, False), ("ifft", True, False), ("coset_fft", False, True), ("coset_ifft", True, True)):
        out = pymodel.ntt(ints, inverse=inv, coset=cos)
        assert digest(oracle.fr_from_ints(out)) == e[name], name
