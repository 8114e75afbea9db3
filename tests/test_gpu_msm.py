"""GPU parity: device MSM (czk_msm_g1 / czk_msm_g2 / czk_msm_bases) vs the CPU oracle.

The oracle restates VariableBaseMSM (algebra/ec/src/msm/variable_base.rs:12-106); the identity
MSM == naive sum is the reference's own (never instantiated) template test
algebra/test-templates/src/msm.rs:16-33.  Compared as affine points, bit-exact.
"""
import numpy as np
import pytest

from helpers import jac_to_affine_ints, make_points

pytestmark = pytest.mark.gpu


def _groups(oracle):
    return {"g1": oracle.G1, "g2": oracle.G2}


def _msm(ctx, G, xy, inf, sc, mont=True):
    fn = ctx.msm_g1 if G.g == "g1" else ctx.msm_g2
    return jac_to_affine_ints(G, fn(xy, inf, sc, mont))


def _oracle_msm(G, xy, inf, sc, mont=True, threads=8):
    out, isinf = G.msm(xy, inf, sc, montgomery=mont, threads=threads)
    return None if isinf else G.affine_to_ints(out)[0]


@pytest.mark.parametrize("g", ["g1", "g2"])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 31, 32, 33, 100, 1000, 4097])
def test_msm_random_matches_oracle(ctx, oracle, g, n):
    G = _groups(oracle)[g]
    xy = make_points(G, max(n, 1), seed=100 + n)[:n]
    sc = oracle.random_fr_mont(200 + n, max(n, 1))[:n]
    inf = np.zeros(n, np.uint8)
    if n > 10:
        inf[::7] = 1
    assert _msm(ctx, G, xy, inf, sc) == _oracle_msm(G, xy, inf, sc)


@pytest.mark.parametrize("g", ["g1", "g2"])
def test_msm_equals_naive_sum(ctx, oracle, g):
    # algebra/test-templates/src/msm.rs:16-33 (there: 2^10 samples; naive is O(n * 256) adds)
    G = _groups(oracle)[g]
    n = 1 << 8
    xy = make_points(G, n, seed=5)
    sc = oracle.random_fr_mont(6, n)
    out, isinf = G.msm_naive(xy, None, sc)
    assert _msm(ctx, G, xy, None, sc) == (None if isinf else G.affine_to_ints(out)[0])


@pytest.mark.parametrize("g", ["g1", "g2"])
def test_msm_edge_scalar_distributions(ctx, oracle, pymodel, g):
    G = _groups(oracle)[g]
    n = 600
    xy = make_points(G, n, seed=11)
    r = pymodel.R_MOD
    cases = {
        "all_zero": [0] * n,
        "all_one": [1] * n,                       # variable_base.rs:44-48 shortcut
        "r_minus_1": [r - 1] * n,
        "small": [(i * 7) % 13 for i in range(n)],  # scalars < 2^c
        "window_edges": [(1 << (16 * (i % 15))) * ((1 << 15) + (i % 3) - 1) % r for i in range(n)],
        "mixed": [0, 1, r - 1, 2, r - 2] * (n // 5),
    }
    for name, vals in cases.items():
        sc = oracle.fr_from_ints(vals)
        assert _msm(ctx, G, xy, None, sc) == _oracle_msm(G, xy, None, sc), name
    # canonical (BigInt256) scalars, the VariableBaseMSM entry
    canon = oracle.random_fr_canonical(3, n)
    assert _msm(ctx, G, xy, None, canon, mont=False) == _oracle_msm(G, xy, None, canon, mont=False)


@pytest.mark.parametrize("g", ["g1", "g2"])
def test_msm_repeated_base_and_cancelling_pairs(ctx, oracle, pymodel, g):
    """Adversarial bucket contents: one repeated base (the criterion bench shape,
    curves/curve-benches/src/macros/ec.rs:199-213: every bucket add is a doubling) and P, -P pairs."""
    G = _groups(oracle)[g]
    n = 512
    one = make_points(G, 1, seed=21)
    xy = np.repeat(one, n, axis=0)
    sc = oracle.random_fr_mont(22, n)
    assert _msm(ctx, G, xy, None, sc) == _oracle_msm(G, xy, None, sc)
    same = np.repeat(oracle.random_fr_mont(23, 1), n, axis=0)
    assert _msm(ctx, G, xy, None, same) == _oracle_msm(G, xy, None, same)
    # P_i = -P_j pairs with equal scalars: the sum is infinity
    pts = make_points(G, n // 2, seed=24)
    ints = G.affine_to_ints(pts)
    neg = pymodel.g1_neg if g == "g1" else pymodel.g2_neg
    both, _ = G.affine_from_ints(ints + [neg(p) for p in ints])
    sc_half = oracle.random_fr_mont(25, n // 2)
    sc2 = np.concatenate([sc_half, sc_half])
    assert _msm(ctx, G, both, None, sc2) is None
    assert _oracle_msm(G, both, None, sc2) is None


def test_msm_uses_min_len(ctx, oracle):
    G = oracle.G1
    xy = make_points(G, 50, seed=31)
    sc = oracle.random_fr_mont(32, 40)
    assert _msm(ctx, G, xy, None, sc) == _oracle_msm(G, xy[:40], None, sc)


@pytest.mark.parametrize("log_n", [16, 18])
def test_msm_g1_large_matches_oracle(ctx, oracle, log_n):
    G = oracle.G1
    n = 1 << log_n
    xy = make_points(G, n, seed=0x377 + log_n, threads=oracle.cpu_threads())
    sc = oracle.random_fr_mont(0x377 + log_n, n)
    inf = np.zeros(n, np.uint8)
    inf[1023::1024] = 1
    assert _msm(ctx, G, xy, inf, sc) == _oracle_msm(G, xy, inf, sc, threads=oracle.cpu_threads())


def test_msm_g2_2_16_matches_oracle(ctx, oracle):
    G = oracle.G2
    n = 1 << 16
    xy = make_points(G, n, seed=0x99, threads=oracle.cpu_threads())
    sc = oracle.random_fr_mont(0x98, n)
    assert _msm(ctx, G, xy, None, sc) == _oracle_msm(G, xy, None, sc, threads=oracle.cpu_threads())


def test_msm_device_resident_bases_and_synthetic(ctx, oracle):
    """czk_bases_synthetic + czk_msm_bases (the resident-CRS path the prover uses) vs the oracle on the
    downloaded bases; also linearity MSM(s) + MSM(t) == MSM(s + t) at 2^20 (size-independent property)."""
    n = 1 << 14
    b = ctx.bases_synthetic(1, seed=7, n=n, inf_every=1024)
    xy, inf = b.numpy()
    assert inf.sum() == n // 1024
    sc = oracle.random_fr_mont(8, n)
    dsc = ctx.vec_from(sc)
    got = jac_to_affine_ints(oracle.G1, ctx.msm_bases(b, dsc))
    assert got == _oracle_msm(oracle.G1, xy, inf, sc)
    # offsets (calculate_coeff uses query[1..], prover.rs:224)
    got = jac_to_affine_ints(oracle.G1, ctx.msm_bases(b, dsc, n=n - 5, base_off=5, sc_off=2))
    assert got == _oracle_msm(oracle.G1, xy[5:], inf[5:], sc[2:n - 3])


def test_msm_linearity_at_2_20(ctx, oracle, pymodel):
    n = 1 << 20
    b = ctx.bases_synthetic(1, seed=9, n=n, inf_every=1024)
    s = oracle.random_fr_mont(10, n)
    t = oracle.random_fr_mont(11, n)
    ds, dt, dst = ctx.vec_from(s), ctx.vec_from(t), ctx.vec_from(oracle.fr_add(s, t))
    G = oracle.G1
    ps = jac_to_affine_ints(G, ctx.msm_bases(b, ds))
    pt = jac_to_affine_ints(G, ctx.msm_bases(b, dt))
    pst = jac_to_affine_ints(G, ctx.msm_bases(b, dst))
    assert pymodel.g1_add(ps, pt) == pst
    assert pymodel.g1_on_curve(pst)


@pytest.mark.parametrize("g", ["g1", "g2"])
def test_msm_merged_window_table_matches_oracle(ctx, oracle, g):
    """czk_bases_precompute: merged-window MSM over the table 2^(c w) P_i must give the same group element,
    including infinity bases, sub-ranges (query[1..]), edge scalars and several window sizes."""
    G = oracle.G1 if g == "g1" else oracle.G2
    n = 1 << 12
    xy = make_points(G, n, seed=61)
    inf = np.zeros(n, np.uint8)
    inf[5::97] = 1
    sc = oracle.random_fr_mont(62, n)
    sc[7] = 0
    sc[8] = oracle.fr_from_ints([1])[0]
    dsc = ctx.vec_from(sc)
    for c in (0, 8, 11, 13, 18):  # 13: bit-slice reduction with 2048 threads per slice; 18: nb / 32 threads
        b = ctx.bases_upload(1 if g == "g1" else 2, xy, inf).precompute(c)
        assert jac_to_affine_ints(G, ctx.msm_bases(b, dsc)) == _oracle_msm(G, xy, inf, sc)
        got = jac_to_affine_ints(G, ctx.msm_bases(b, dsc, n=n - 3, base_off=3, sc_off=1))
        assert got == _oracle_msm(G, xy[3:], inf[3:], sc[1:n - 2])
        b.free()


@pytest.mark.parametrize("tree", [False, True])
def test_msm_bases_multi_shares_the_digit_sort(ctx, oracle, tree):
    """czk_msm_bases_multi == the oracle for every base set: two G1 sets and a G2 set with the lead's infinity flags
    (one digit sort serves all three, prover.rs:104-160), a set with other flags in the middle (it and everything
    after it run as MSMs of their own), sets without a table, and the query[1..] sub-range the prover uses.  Both
    accumulation algorithms (the small input takes the affine tree only when told to)."""
    n = 1 << 12
    sets = [(oracle.G1, 1, 81), (oracle.G1, 1, 82), (oracle.G2, 2, 83), (oracle.G1, 1, 84)]
    inf = np.zeros(n, np.uint8)
    inf[3::61] = 1
    other = np.zeros(n, np.uint8)
    other[4::61] = 1
    sc = oracle.random_fr_mont(85, n)
    dsc = ctx.vec_from(sc)
    xys = [make_points(G, n, seed=seed) for G, _, seed in sets]

    def check(flags, tables, base_off=0):
        bs = [ctx.bases_upload(curve, xy, f) for (_, curve, _), xy, f in zip(sets, xys, flags)]
        for b, t in zip(bs, tables):
            if t:
                b.precompute(11)
        m = n - base_off
        outs = ctx.msm_bases_multi(bs, dsc, n=m, base_off=base_off)
        for (G, _, _), xy, f, out in zip(sets, xys, flags, outs):
            assert jac_to_affine_ints(G, out) == _oracle_msm(G, xy[base_off:], f[base_off:], sc[:m])
        for b in bs:
            b.free()

    try:
        ctx.msm_set_batched(True, always=tree)
        check([inf, inf, inf, inf], [True] * 4)
        check([inf, inf, inf, inf], [True] * 4, base_off=1)
        check([inf, other, inf, inf], [True] * 4)
        check([inf, inf, inf, other], [True, True, True, False])
        check([inf, inf, inf, inf], [False] * 4)
    finally:
        ctx.msm_set_batched(True)


def test_msm_merged_large_and_linear(ctx, oracle, pymodel):
    n = 1 << 18
    b = ctx.bases_synthetic(1, seed=71, n=n, inf_every=1024)
    xy, inf = b.numpy()
    sc = oracle.random_fr_mont(72, n)
    dsc = ctx.vec_from(sc)
    plain = jac_to_affine_ints(oracle.G1, ctx.msm_bases(b, dsc))
    b.precompute()
    merged = jac_to_affine_ints(oracle.G1, ctx.msm_bases(b, dsc))
    assert merged == plain == _oracle_msm(oracle.G1, xy, inf, sc, threads=oracle.cpu_threads())


def test_fq_inverse_device_matches_oracle(ctx, oracle, pymodel):
    """The one inversion per block of the batched-affine accumulation: binary extended Euclid on the device
    (algebra/ff/src/fields/macros.rs:368-422) vs the oracle's restatement, edge values included."""
    q = pymodel.Q_MOD
    vals = [1, 2, q - 1, q - 2, (q + 1) // 2, 3, 1 << 376, (1 << 376) + 1] + [pow(7, 1000 + i, q) for i in range(120)]
    a = oracle.fq_from_ints(vals)
    assert (ctx.fq_inverse(a) == oracle.fq_inv(a)).all()
    z = oracle.fq_from_ints([0])
    assert (ctx.fq_inverse(z) == z).all()


@pytest.mark.parametrize("g", ["g1", "g2"])
def test_msm_batched_affine_equals_xyzz_walk(ctx, oracle, g):
    """Both accumulation algorithms give the same group element (and both match the oracle): random scalars (the
    affine tree runs to the end), a repeated base and cancelling pairs (it must raise its flag and hand over)."""
    G = _groups(oracle)[g]
    n = 3000
    xy = make_points(G, n, seed=71)
    sc = oracle.random_fr_mont(72, n)
    exp = _oracle_msm(G, xy, None, sc)
    try:
        ctx.msm_set_batched(False)
        l0 = ctx.launches
        walk = _msm(ctx, G, xy, None, sc)
        l1 = ctx.launches
        ctx.msm_set_batched(True, always=True)  # small inputs: the tree would not be chosen by size
        tree = _msm(ctx, G, xy, None, sc)
        assert walk == exp and tree == exp
        assert ctx.launches - l1 > l1 - l0, "the round kernels of the affine tree did not run"
        rep = np.repeat(xy[:1], 700, axis=0)
        sc7 = oracle.random_fr_mont(73, 700)
        assert _msm(ctx, G, rep, None, sc7) == _oracle_msm(G, rep, None, sc7)
        # skewed bucket loads: few distinct scalars => some very long buckets next to empty ones (many rounds)
        few = np.tile(oracle.random_fr_mont(74, 3), (1000, 1))
        assert _msm(ctx, G, xy, None, few) == _oracle_msm(G, xy, None, few)
    finally:
        ctx.msm_set_batched(True)
