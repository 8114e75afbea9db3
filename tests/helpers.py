"""Shared helpers for the parity tests (test infrastructure; may use oracle/)."""
import numpy as np

from oracle import binding as o
from oracle import pymodel as m


def jac_to_affine_ints(G, xyz):
    """Product MSM output (x | y | z limbs, z in {0, 1 (Montgomery)}) -> None | affine ints."""
    w = G.w
    half = w // 2
    xyz = np.asarray(xyz, dtype=np.uint64)
    z = xyz[w:w + half]
    if not z.any():
        return None
    one = o.fq_from_ints([1])[0]
    assert (z[:6] == one).all() and not z[6:].any(), "MSM output must be affine-normalised (z = 1)"
    return G.affine_to_ints(xyz[:w])[0]


def make_points(G, n, seed, threads=8):
    """n distinct points of the prime-order subgroup: (k0 + i*kstep) * generator."""
    gen = m.G1_GEN if G.g == "g1" else m.G2_GEN
    gxy, _ = G.affine_from_ints([gen])
    ks = o.random_fr_mont(seed, 2)
    return G.gen_progression(gxy[0], ks[0], ks[1], n, threads=threads)
