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


def plonk_wiring_instance(log_d, seed, size=None):
    """A VALID wiring instance over the domain of size 2^log_d (or of `size` = 3 * 2^k points, the reference's mixed-radix
    wire domain): a permutation sigma made of short cycles, wire values that
    are constant on every cycle, p = interpolate(values), w = interpolate(omega^sigma(i)) - so that
    prod_i (p_i + y w_i + z) / (p_i + y omega^i + z) = 1 and the unit-product argument has something true to prove.
    Returns (p_coeffs, w_coeffs) as Montgomery Fr arrays."""
    import random

    D = size if size is not None else 1 << log_d
    rnd = random.Random(seed)
    group_gen = o.mixed_domain_params(D)[0]  # == Radix2EvaluationDomain's for a power of two
    omega_pows = np.zeros((D, 4), np.uint64)
    cur = o.fr_from_ints([1])[0]
    for i in range(D):
        omega_pows[i] = cur
        cur = o.fr_mul(cur[None, :], group_gen[None, :])[0]
    idx = list(range(D))
    rnd.shuffle(idx)
    sigma = list(range(D))
    vals = [0] * D
    pos = 0
    while pos < D:
        ln = min(rnd.choice((1, 2, 3, 5)), D - pos)
        cyc = idx[pos:pos + ln]
        v = rnd.randrange(m.R_MOD)
        for a, b in zip(cyc, cyc[1:] + cyc[:1]):
            sigma[a] = b
            vals[a] = v
        pos += ln
    p_evals = o.fr_from_ints(vals)
    w_evals = omega_pows[sigma]
    if D & (D - 1):
        return o.ntt_mixed(p_evals, inverse=True), o.ntt_mixed(np.ascontiguousarray(w_evals), inverse=True)
    return o.ntt(p_evals, inverse=True), o.ntt(np.ascontiguousarray(w_evals), inverse=True)


def kzg_powers(n, tau_int, threads=8):
    """powers_of_g = tau^i * G1 generator, i < n (the KZG10 committer key), as affine Montgomery limbs."""
    g1, _ = o.generators()
    out = np.zeros((n, 12), np.uint64)
    t = 1
    for i in range(n):
        out[i], inf = o.G1.scalar_mul(g1, o.fr_from_ints([t])[0])
        assert not inf
        t = t * tau_int % m.R_MOD
    return out
