"""Frozen known-answer vectors (tests/golden/vectors.json, written by tests/golden/make_vectors.py): the oracle must keep
reproducing them (CPU tier), and the device path must hit the same values (GPU tier)."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from helpers import jac_to_affine_ints, kzg_powers, make_points, plonk_wiring_instance

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


# ---------------------------------------------------------------------------------------------------------------------
# mixed-radix transforms (3 * 2^k points) and the Plonk wiring argument
_FLAVOURS = (("fft", False, False), ("ifft", True, False), ("coset_fft", False, True), ("coset_ifft", True, True))


@pytest.mark.parametrize("e", VEC["ntt_mixed"], ids=lambda e: f"3*2^{e['log_m']}")
def test_oracle_ntt_mixed_golden(oracle, e):
    v = oracle.random_fr_mont(e["seed"], 3 << e["log_m"])
    for name, inv, cos in _FLAVOURS:
        assert digest(oracle.ntt_mixed(v, inv, cos)) == e[name], name


@pytest.mark.parametrize("e", VEC["ntt_mixed"], ids=lambda e: f"3*2^{e['log_m']}")
def test_plain_dft_ntt_mixed_golden(oracle, pymodel, e):
    """The frozen digests are the plain O(n^2) DFT under get_root_of_unity(n) in Python integers (no shared code with the C
    restatement of the reference's permute / radix-3 / radix-2 passes)."""
    R = pymodel.R_MOD
    n = 3 << e["log_m"]
    x = oracle.fr_to_ints(oracle.random_fr_mont(e["seed"], n))
    w, w_inv, n_inv, g_inv = [oracle.fr_to_ints([t])[0] for t in oracle.mixed_domain_params(n)]
    g = pow(g_inv, R - 2, R)

    def dft(vals, root):
        pw = [1] * n
        for i in range(1, n):
            pw[i] = pw[i - 1] * root % R
        return [sum(vals[i] * pw[(i * j) % n] for i in range(n)) % R for j in range(n)]

    gp = [pow(g, i, R) for i in range(n)]
    gip = [pow(g_inv, i, R) for i in range(n)]
    want = {"fft": dft(x, w), "ifft": [v * n_inv % R for v in dft(x, w_inv)],
            "coset_fft": dft([a * b % R for a, b in zip(x, gp)], w),
            "coset_ifft": [v * n_inv % R * b % R for v, b in zip(dft(x, w_inv), gip)]}
    for name, _, _ in _FLAVOURS:
        assert digest(oracle.fr_from_ints(want[name])) == e[name], name


@pytest.mark.gpu
@pytest.mark.parametrize("e", VEC["ntt_mixed"], ids=lambda e: f"3*2^{e['log_m']}")
def test_device_ntt_mixed_golden(ctx, oracle, e):
    v = oracle.random_fr_mont(e["seed"], 3 << e["log_m"])
    for name, inv, cos in _FLAVOURS:
        assert digest(ctx.ntt_mixed(v, inv, cos)) == e[name], name


def _wiring_inputs(oracle, e):
    powers = kzg_powers(e["D"], e["tau"])
    p, w = plonk_wiring_instance(None, seed=e["instance_seed"], size=e["D"])
    return powers, p, w


@pytest.mark.parametrize("e", VEC["plonk_wiring"], ids=lambda e: f"{e['scheme']}-{e['parties']}p-D{e['D']}")
def test_oracle_plonk_wiring_golden(oracle, e):
    powers, p, w = _wiring_inputs(oracle, e)
    shares = p[None] if e["parties"] == 1 else oracle.king_share_batch(p, e["parties"], seed=9)
    scheme = oracle.SCHEME_PLAIN if e["scheme"] == "plain" else oracle.SCHEME_SPDZ
    res = oracle.plonk_prove_wiring(scheme, shares, w, powers, seed=e["transcript_seed"], threads=2)
    assert res["status"] == 1
    pf = res["proof"]
    got = dict(cmt=digest(pf["cmt_xy"]), open_val=digest(pf["open_val"]), open_pf=digest(pf["open_pf_xy"]),
               challenges=digest(pf["challenges"]), share_pf=digest(res["share_pf_xy"]))
    assert got == {k: e[k] for k in got}


@pytest.mark.gpu
@pytest.mark.parametrize("e", [e for e in VEC["plonk_wiring"] if e["parties"] == 1], ids=lambda e: f"{e['scheme']}-D{e['D']}")
def test_device_plonk_wiring_golden(ctx, czk, oracle, e):
    ctx.net_init(0, 1, None)
    powers, p, w = _wiring_inputs(oracle, e)
    D = e["D"]
    mixed = bool(D & (D - 1))
    log = (D // 3 if mixed else D).bit_length() - 1
    b = ctx.bases_upload(1, powers)
    try:
        got = czk.plonk_prove_wiring(ctx, czk.SCHEME_PLAIN, b, log, ctx.vec_from(p), None, ctx.vec_from(w), seed=e["transcript_seed"], mixed=mixed)
    finally:
        b.free()
    pf = got["proof"]
    assert digest(pf["cmt_xy"]) == e["cmt"] and digest(pf["open_val"]) == e["open_val"]
    assert digest(pf["open_pf_xy"]) == e["open_pf"] and digest(pf["challenges"]) == e["challenges"]
