#!/usr/bin/env python3
"""Freeze known-answer vectors of the oracle for the hot path into tests/golden/vectors.json.

TEST INFRASTRUCTURE.  The reference holds no byte vectors for MSM, NTT or proofs, and it cannot be run here
(Rust), so these are NOT reference outputs: they are the oracle's outputs at the commit where it passed every identity the
reference's tests state (tests/test_oracle*.py, tests/test_verifier.py).  Committing them pins the oracle against silent
drift and gives the CPU tier and the GPU tier a third, frozen point of comparison.  Inputs come from the seeded generators
of oracle/binding.py, so the file stays small: seeds in, canonical integers (hex) out.

    python tests/golden/make_vectors.py        # rewrites tests/golden/vectors.json
"""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np

from oracle import binding as o
from helpers import kzg_powers, make_points, plonk_wiring_instance


def digest(arr) -> str:
    return hashlib.sha256(np.ascontiguousarray(arr, np.uint64).tobytes()).hexdigest()


def hexpt(G, xy, inf=0):
    if inf:
        return None
    p = G.affine_to_ints(np.ascontiguousarray(xy).reshape(1, -1))[0]
    flat = p if G.g == "g1" else (p[0][0], p[0][1], p[1][0], p[1][1])
    return [hex(v) for v in flat]


def main():
    o.build()
    out = {"about": "oracle outputs frozen by tests/golden/make_vectors.py; inputs are the seeded generators named in each entry"}
    # MSM: bases make_points(G, n, seed), scalars random_fr_mont(seed, n)
    out["msm"] = []
    for g, n, seed in (("g1", 1000, 11), ("g1", 4097, 12), ("g2", 257, 13)):
        G = o.G1 if g == "g1" else o.G2
        xy = make_points(G, n, seed=seed)
        sc = o.random_fr_mont(seed + 100, n)
        inf = np.zeros(n, np.uint8)
        inf[::7] = 1
        res, isinf = G.msm(xy, inf, sc, threads=4)
        out["msm"].append(dict(group=g, n=n, points_seed=seed, scalars_seed=seed + 100, inf_every=7, result=hexpt(G, res, isinf)))
    # NTT: random_fr_mont(seed, 2^k) through the four transforms: sha256 of the little-endian u64 Montgomery limbs
    out["ntt"] = []
    for log_d, seed in ((10, 21), (13, 22)):
        v = o.random_fr_mont(seed, 1 << log_d)
        e = dict(log_d=log_d, seed=seed)
        for name, inv, cos in (("fft", False, False), ("ifft", True, False), ("coset_fft", False, True), ("coset_ifft", True, True)):
            e[name] = digest(o.ntt(v, inv, cos))
        out["ntt"].append(e)
    # Groth16: toxic = random_fr_mont(31, 7), chain from random_fr_mont(32, 1), r = random_fr_mont(33, 1), s = random_fr_mont(34, 1)
    out["groth16"] = []
    for n_sq in (10, 1 << 10):
        toxic = o.random_fr_mont(31, 7)
        pk = o.groth16_setup(n_sq, toxic, threads=4)
        chain = o.squaring_chain(o.random_fr_mont(32, 1)[0], n_sq)
        r, s = o.random_fr_mont(33, 1), o.random_fr_mont(34, 1)
        res = o.groth16_prove(o.SCHEME_PLAIN, n_sq, [chain], r, s, pk, threads=4)
        assert res["ok"]
        out["groth16"].append(dict(n_sq=n_sq, proof=digest(res["proof"]), h=digest(res["h"][0]),
                                   a=hexpt(o.G1, res["proof"][:12]), c=hexpt(o.G1, res["proof"][36:48]),
                                   pk_digest=digest(np.concatenate([pk[k].reshape(-1) for k in ("a_query", "b_g1_query", "b_g2_query", "h_query", "l_query")]))))
    # mixed-radix NTT (3 * 2^log_m points): random_fr_mont(seed, 3 << log_m) through the four transforms
    out["ntt_mixed"] = []
    for log_m, seed in ((5, 41), (8, 42)):
        v = o.random_fr_mont(seed, 3 << log_m)
        e = dict(log_m=log_m, seed=seed)
        for name, inv, cos in (("fft", False, False), ("ifft", True, False), ("coset_fft", False, True), ("coset_ifft", True, True)):
            e[name] = digest(o.ntt_mixed(v, inv, cos))
        out["ntt_mixed"].append(e)
    # Plonk wiring argument: kzg_powers(D, tau), plonk_wiring_instance(size=D, seed), stand-in transcript seed 77;
    # n-party entries share p with king_share_batch(p, parties, seed=9)
    out["plonk_wiring"] = []
    for D, parties, scheme in ((64, 1, "plain"), (48, 1, "plain"), (48, 2, "spdz")):
        tau, seed = 0x60 + D, 50 + D
        powers = kzg_powers(D, tau)
        p, w = plonk_wiring_instance(None, seed=seed, size=D)
        shares = p[None] if parties == 1 else o.king_share_batch(p, parties, seed=9)
        res = o.plonk_prove_wiring(o.SCHEME_PLAIN if scheme == "plain" else o.SCHEME_SPDZ, shares, w, powers, seed=77, threads=2)
        assert res["status"] == 1
        pf = res["proof"]
        out["plonk_wiring"].append(dict(D=D, parties=parties, scheme=scheme, tau=tau, instance_seed=seed, transcript_seed=77,
                                        cmt=digest(pf["cmt_xy"]), open_val=digest(pf["open_val"]), open_pf=digest(pf["open_pf_xy"]),
                                        challenges=digest(pf["challenges"]), share_pf=digest(res["share_pf_xy"])))
    (Path(__file__).resolve().parent / "vectors.json").write_text(json.dumps(out, indent=1))
    print("wrote tests/golden/vectors.json")


if __name__ == "__main__":
    main()
