"""The DEVICE field / curve headers (collaborative-zksnark_b200/csrc/{carry,fp,ec,msm_digits}.cuh),
compiled for the host with the PTX carry flag emulated (tests/emu), against the oracle.  This checks the
exact limb schedule the GPU executes without needing a GPU; the -m gpu tests check the kernels."""
import ctypes as C
import random
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-C", str(ROOT / "tests/emu"), "-s"], check=True)
    return C.CDLL(str(ROOT / "tests/emu/libczk_emu.so"))


def P(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _edge(vals_list, mod):
    for v in vals_list:
        v[0:4] = [0, mod - 1, 1, mod - 1]
    if len(vals_list) > 1:
        vals_list[1][0:4] = [0, mod - 1, mod - 1, 1]
        vals_list[0][5] = vals_list[1][5]  # a - a


@pytest.mark.parametrize("name,nl,nargs", [
    ("fr_mul", 4, 2), ("fr_add", 4, 2), ("fr_sub", 4, 2), ("fr_neg", 4, 1),
    ("fq_mul", 6, 2), ("fq_add", 6, 2), ("fq_sub", 6, 2),
])
def test_prime_field_ops(emu, oracle, pymodel, name, nl, nargs):
    rnd = random.Random(hash(name) & 0xffff)
    mod = pymodel.R_MOD if nl == 4 else pymodel.Q_MOD
    n = 3000
    vals = [[rnd.randrange(mod) for _ in range(n)] for _ in range(nargs)]
    _edge(vals, mod)
    frm = oracle.fr_from_ints if nl == 4 else oracle.fq_from_ints
    arrs = [frm(v) for v in vals]
    out = np.zeros_like(arrs[0])
    getattr(emu, "emu_" + name)(P(out), *[P(a) for a in arrs], C.c_size_t(n))
    assert (out == getattr(oracle, name)(*arrs)).all()


def test_fq2_and_inversions(emu, oracle, pymodel):
    rnd = random.Random(7)
    a = oracle.fq_from_ints([rnd.randrange(pymodel.Q_MOD) for _ in range(400)]).reshape(-1, 12)
    b = oracle.fq_from_ints([rnd.randrange(pymodel.Q_MOD) for _ in range(400)]).reshape(-1, 12)
    out = np.zeros_like(a)
    emu.emu_fq2_mul(P(out), P(a), P(b), C.c_size_t(200))
    assert (out == oracle.fq2_mul(a, b)).all()
    emu.emu_fq2_sqr(P(out), P(a), C.c_size_t(200))
    assert (out == oracle.fq2_sqr(a)).all()
    emu.emu_fq2_inv(P(out[:10]), P(a[:10]), C.c_size_t(10))
    assert (out[:10] == oracle.fq2_inv(a[:10])).all()
    x = oracle.fq_from_ints([rnd.randrange(1, pymodel.Q_MOD) for _ in range(10)])
    o6 = np.zeros_like(x)
    emu.emu_fq_inv(P(o6), P(x), C.c_size_t(10))
    assert (o6 == oracle.fq_inv(x)).all()
    y = oracle.fr_from_ints([rnd.randrange(1, pymodel.R_MOD) for _ in range(10)])
    o4 = np.zeros_like(y)
    emu.emu_fr_inv(P(o4), P(y), C.c_size_t(10))
    assert (o4 == oracle.fr_inv(y)).all()
    emu.emu_fr_from_mont(P(o4), P(y), C.c_size_t(10))
    assert (o4 == oracle.fr_into_repr(y)).all()


def test_interleaved_double_product(emu, oracle, pymodel):
    """Fp::mul2 (two products with interleaved rows) == two products; aliasing an output with an input as the microbenchmark does."""
    rnd = random.Random(13)
    a, b, c, d = (oracle.fq_from_ints([rnd.randrange(pymodel.Q_MOD) for _ in range(300)]) for _ in range(4))
    r1, r2 = np.zeros_like(a), np.zeros_like(a)
    emu.emu_fq_mul2(P(r1), P(r2), P(a), P(b), P(c), P(d), C.c_size_t(300))
    assert (r1 == oracle.fq_mul(a, b)).all() and (r2 == oracle.fq_mul(c, d)).all()


def test_batched_step_binary_gcd_inversion(emu, oracle, pymodel):
    """csrc/fq_inverse.cuh (Pornin's batched binary GCD, 31 steps per round on 64-bit approximations) against the oracle's
    restatement of the reference's inverse (macros.rs:368-422): random values, the values that need every round (powers of
    two, p - 1, small numbers whose partner is the full-length modulus) and values that fit one or two limbs."""
    rnd = random.Random(11)
    for mod, from_ints, inv, fn, nl in ((pymodel.Q_MOD, oracle.fq_from_ints, oracle.fq_inv, emu.emu_fq_inv_bingcd, 6),
                                        (pymodel.R_MOD, oracle.fr_from_ints, oracle.fr_inv, emu.emu_fr_inv_bingcd, 4)):
        bits = mod.bit_length()
        vals = [1, 2, 3, mod - 1, mod - 2, (mod + 1) // 2, 1 << (bits - 1), (1 << (bits - 1)) + 1, 1 << 31, (1 << 31) - 1, 1 << 32,
                (1 << 62) + 1, (1 << 64) - 1, 1 << 64, (1 << 95) + 12345]
        vals += [rnd.randrange(1, mod) for _ in range(1500)]
        vals += [rnd.randrange(1, 1 << k) for k in (1, 5, 31, 33, 63, 64, 65, 100, 200) for _ in range(30)]
        # the device sees Montgomery-form inputs: take these integers AS the Montgomery limbs too (any value < p is one)
        x = from_ints(vals)
        out = np.zeros_like(x)
        fn(P(out), P(x), C.c_size_t(len(vals)))
        assert (out == inv(x)).all()
        raw = np.array([[(v >> (64 * i)) & (2**64 - 1) for i in range(nl)] for v in vals], dtype=np.uint64)
        fn(P(out), P(raw), C.c_size_t(len(vals)))
        assert (out == inv(raw)).all()


@pytest.mark.parametrize("g", ["g1", "g2"])
def test_xyzz_bucket_accumulation(emu, oracle, pymodel, g):
    """Signed accumulation exactly like a device bucket, incl. P+P (doubling branch), P + -P, and the
    XYZZ+XYZZ path used by the bucket reduction."""
    G = oracle.G1 if g == "g1" else oracle.G2
    gen = pymodel.G1_GEN if g == "g1" else pymodel.G2_GEN
    mul, add, neg = (pymodel.g1_mul, pymodel.g1_add, pymodel.g1_neg) if g == "g1" else (pymodel.g2_mul, pymodel.g2_add, pymodel.g2_neg)
    fn = emu.emu_g1_sum if g == "g1" else emu.emu_g2_sum
    ks = [3, 3, 5, 7, 7, 7, 11, 13, 2, 9, 9, 4]
    pts = [mul(gen, k) for k in ks]
    sign = np.array([0, 0, 0, 0, 1, 0, 1, 0, 0, 1, 0, 0], np.uint8)
    expect = None
    for p, s in zip(pts, sign):
        expect = add(expect, neg(p) if s else p)
    xy, _ = G.affine_from_ints(pts)
    out = np.zeros(G.w, np.uint64)
    for tree in (0, 1):
        isinf = fn(P(out), P(xy), sign.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_size_t(len(pts)), C.c_int(tree))
        assert not isinf and G.affine_to_ints(out)[0] == expect
    xy, _ = G.affine_from_ints([mul(gen, 5), mul(gen, 5), neg(mul(gen, 10))])
    for tree in (0, 1):
        assert fn(P(out), P(xy), None, C.c_size_t(3), C.c_int(tree)) == 1


@pytest.mark.parametrize("c", [2, 3, 8, 11, 13, 15, 16, 17, 23])
def test_signed_digit_decomposition(emu, oracle, pymodel, c):
    rnd = random.Random(c)
    nwin = (254 + c - 1) // c
    cases = [0, 1, pymodel.R_MOD - 1, (1 << 252) + 12345, 1 << (c - 1), (1 << c) - 1, (1 << (c - 1)) + 1]
    cases += [rnd.randrange(pymodel.R_MOD) for _ in range(300)]
    for s in cases:
        out = np.zeros(nwin, np.int32)
        emu.emu_signed_digits(out.ctypes.data_as(C.POINTER(C.c_int32)), P(oracle.ints_to_limbs([s], 4)), C.c_uint(c), C.c_uint(nwin))
        assert sum(int(d) << (c * w) for w, d in enumerate(out)) == s
        assert all(abs(int(d)) <= 1 << (c - 1) for d in out)


def test_lazy_reduction_sum_of_two_products(emu, oracle, pymodel):
    """Fp::mul_sum2 = (a b + c d) / R with one Montgomery reduction, and the two-lane Fq2 product built from it (the G2
    bucket kernel: lane 0 computes c0 = a0 b0 - 5 a1 b1 with -5 a1 passed as the UNREDUCED integer 5 (p - a1), lane 1
    computes c1 = a1 b0 + a0 b1), against the oracle's Karatsuba Fq2 product (quadratic_extension.rs:569-583)."""
    rnd = random.Random(17)
    q = pymodel.Q_MOD
    n = 600
    ints = [[rnd.randrange(q) for _ in range(2 * n)] for _ in range(2)]
    # extremes: zero / p - 1 components in every combination (5 (p - 0) = 5 p is the largest unreduced operand)
    edge = [0, q - 1, 1, q - 2]
    k = 0
    for u in edge:
        for v in edge:
            for w in edge:
                ints[0][2 * k], ints[0][2 * k + 1], ints[1][2 * k], ints[1][2 * k + 1] = u, v, w, edge[(k * 7) % 4]
                k += 1
    a = oracle.fq_from_ints(ints[0]).reshape(-1, 12)
    b = oracle.fq_from_ints(ints[1]).reshape(-1, 12)
    # the same integers taken AS Montgomery limbs as well (any value < p is a valid element): covers limbs like 0 and p - 1
    for x, y in ((a, b), (np.array([[(v >> (64 * i)) & (2**64 - 1) for v in pair for i in range(6)] for pair in zip(ints[0][::2], ints[0][1::2])], np.uint64),
                          np.array([[(v >> (64 * i)) & (2**64 - 1) for v in pair for i in range(6)] for pair in zip(ints[1][::2], ints[1][1::2])], np.uint64))):
        out = np.zeros_like(x)
        emu.emu_fq2_mul_two_lanes(P(out), P(x), P(y), C.c_size_t(x.shape[0]))
        assert (out == oracle.fq2_mul(x, y)).all()
    r = pymodel.R_MOD
    fa, fb, fc, fd = (oracle.fr_from_ints([rnd.randrange(r) for _ in range(500)]) for _ in range(4))
    fo = np.zeros_like(fa)
    emu.emu_fr_mul_sum2(P(fo), P(fa), P(fb), P(fc), P(fd), C.c_size_t(500))
    assert (fo == oracle.fr_add(oracle.fr_mul(fa, fb), oracle.fr_mul(fc, fd))).all()


@pytest.mark.parametrize("n,tile_log", [(3, 11), (4, 11), (5, 11), (7, 11), (10, 11), (11, 11), (12, 11), (13, 11), (14, 11),
                                        (9, 5), (10, 6), (11, 6), (12, 6), (13, 7), (15, 6), (16, 7)])
def test_ntt_tile_phases_match_oracle(emu, oracle, n, tile_log):
    """csrc/ntt_tile.cuh - the register-phase NTT the device runs (DIF and DIT orders, inverse twiddles as -tw[D/2 - e],
    unit-twiddle shortcuts, fused scalings, the pass plan) - executed thread by thread on the host for all five
    transforms and compared with the oracle's restatement of radix2/fft.rs.  Small tiles force multi-pass plans (up to
    4 passes) at sizes the CPU finishes quickly; tile_log 11 is the device's own plan."""
    d = 1 << n
    dp = oracle.domain_params(d)
    gen22 = oracle.fr_from_ints([22])[0]  # the coset shift: Fr::multiplicative_generator() (fr.rs, GENERATOR)
    x = oracle.random_fr_mont(0x47 + n, d)
    f = emu.emu_ntt
    f.restype = C.c_int
    for op in range(6):  # 4: the fused pair with two-level scale tables, 5: with the pre-permuted one-factor-per-position table
        got = x.copy()
        assert f(P(got), n, op, tile_log, P(dp["group_gen"]), P(gen22), P(dp["generator_inv"]), P(dp["size_inv"])) == 1
        if op < 4:
            exp = oracle.ntt(x, inverse=bool(op & 1), coset=bool(op & 2))
        else:
            exp = oracle.ntt(oracle.ntt(x, inverse=True, coset=False), inverse=False, coset=True)
        assert (got == exp).all(), (n, tile_log, op)


@pytest.mark.parametrize("top", [12, 13, 16])
def test_two_level_grid_reduction_index_maps(top):
    """The index maps of k_msm_grid_sums / k_msm_grid_slices (csrc/msm.cu), restated in Python over integers standing in for
    points: with v = hi * L + lo, the row sums R_hi and column sums C_lo give every bit-slice sum S_j, and
    sum_j 2^j S_j = sum_v v B_v (what the host tail's Horner recombination computes).  The kernels themselves are covered
    bit for bit by the GPU parity tests; this pins the arithmetic identity they rely on, for even and odd bit counts."""
    import random

    rnd = random.Random(top)
    nb = 1 << top
    lo_bits = top // 2
    hi_bits = top - lo_bits
    L, H = 1 << lo_bits, 1 << hi_bits
    B = [rnd.randrange(1 << 40) for _ in range(nb)]  # B[v - 1] is the bucket of weight v, v = 1 .. nb
    R = [sum(B[h * L + k - 1] for k in range(L) if h * L + k) for h in range(H)]
    C = [sum(B[k * L + l - 1] for k in range(H) if k * L + l) for l in range(L)]
    S = []
    for j in range(top + 1):
        if j == top:
            S.append(B[H * L - 1])
            continue
        col = j < lo_bits
        bit = j if col else j - lo_bits
        src, count = (C, L // 2) if col else (R, H // 2)
        S.append(sum(src[((((k >> bit) << 1) | 1) << bit) | (k & ((1 << bit) - 1))] for k in range(count)))
    assert sum(s << j for j, s in enumerate(S)) == sum(v * B[v - 1] for v in range(1, nb + 1))


@pytest.mark.parametrize("n,tile_log", [(4, 4), (7, 5), (9, 11)])
@pytest.mark.parametrize("inverse", [0, 1])
def test_mixed_radix_transform_emulated(emu, oracle, pymodel, n, tile_log, inverse):
    """The device's mixed-radix transform over 3 * 2^n points on the host: de-interleave, the tile kernels' own phase code for
    the three radix-2 transforms, and ntt_mixed_combine3 (the arithmetic of k_mr_combine) - against the oracle's restatement
    of MixedRadixEvaluationDomain."""
    M, N = 1 << n, 3 << n
    dp = oracle.domain_params(M)
    gen22 = oracle.fr_from_ints([22])[0]
    w3, w3_inv, _, _ = oracle.mixed_domain_params(N)
    third_inv = oracle.fr_from_ints([pow(3, pymodel.R_MOD - 2, pymodel.R_MOD)])[0]
    x = oracle.random_fr_mont(0x61 + n, N)
    got = x.copy()
    f = emu.emu_ntt_mixed
    f.restype = C.c_int
    assert f(P(got), n, inverse, tile_log, P(dp["group_gen"]), P(gen22), P(dp["generator_inv"]), P(dp["size_inv"]),
             P(w3_inv if inverse else w3), P(third_inv)) == 1
    assert (got == oracle.ntt_mixed(x, inverse=bool(inverse))).all()
