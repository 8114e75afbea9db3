"""Python big-int model of the BLS12-377 hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.  It is an *independent* second model (textbook affine arithmetic
on Python ints, O(n^2) DFT by direct evaluation) used to pin the C restatement
in oracle/czk_oracle.c, which is the line-by-line one.

What it models, and where the reference defines it (paths relative to
/root/reference):
  * Fr / Fq / Fq2 values                curves/bls12_377/src/fields/{fr,fq,fq2}.rs
  * Montgomery representation           algebra/ff/src/fields/macros.rs:444-454 (from_repr),
                                        algebra/ff/src/fields/arithmetic.rs:59-82 (into_repr)
  * root of unity for a radix-2 domain  algebra/ff/src/fields/mod.rs:337-386 (large-subgroup branch)
  * domain constants                    algebra/poly/src/domain/radix2/mod.rs:51-82
  * FFT semantics out[i] = p(w^i)       algebra/poly/src/domain/radix2/mod.rs:320-360 (the reference's own test)
  * coset FFT = distribute_powers(g)    algebra/poly/src/domain/mod.rs:139-142, g = 22
  * short-Weierstrass group law         algebra/ec/src/models/short_weierstrass_jacobian.rs (affine meaning)
  * MSM meaning sum s_i P_i             algebra/test-templates/src/msm.rs:16-33 (naive reference)
"""
from __future__ import annotations

# ----------------------------------------------------------------------------
# moduli (decimal; re-checked against the golden JSON in tests)
R_MOD = 8444461749428370424248824938781546531375899335154063827935233455917409239041
Q_MOD = 258664426012969094010652733694893533536393512754914660539884262666720468348340822774968888139573360124440321458177

FR_LIMBS = 4
FQ_LIMBS = 6
FR_R = pow(2, 256, R_MOD)
FQ_R = pow(2, 384, Q_MOD)
FR_RINV = pow(FR_R, -1, R_MOD)
FQ_RINV = pow(FQ_R, -1, Q_MOD)

FR_TWO_ADICITY = 47
FR_GENERATOR = 22  # multiplicative generator == coset shift
FR_SMALL_SUBGROUP_BASE = 3
FR_SMALL_SUBGROUP_BASE_ADICITY = 1
# LARGE_SUBGROUP_ROOT_OF_UNITY = 11^((r-1)/(3*2^47))   (decoded from fr.rs:23-28)
FR_LARGE_SUBGROUP_ROOT = pow(11, (R_MOD - 1) // (3 * 2**47), R_MOD)
FR_TWO_ADIC_ROOT = pow(22, (R_MOD - 1) // 2**47, R_MOD)

FQ2_NONRESIDUE = Q_MOD - 5
G1_B = 1
G1_GEN = (
    81937999373150964239938255573465948239988671502647976594219695644855304257327692006745978603320413799295628339695,
    241266749859715473739788878240585681733927191168601896383759122102112907357779751001206799952863815012735208165030,
)
G2_B = (0, 155198655607781456406391640216936120121836107652948796323930557600032281009004493664981332883744016074664192874906)
G2_GEN = (
    (233578398248691099356572568220835526895379068987715365179118596935057653620464273615301663571204657964920925606294,
     140913150380207355837477652521042157274541796891053068589147167627541651775299824604154852141315666357241556069118),
    (63160294768292073209381361943935198908131692476676907196754037919244929611450776219210369229519898517858833747423,
     149157405641012693445398062341192467754805999074082136895788947234480009303640899064710353187729182149407503257491),
)


# ----------------------------------------------------------------------------
# Montgomery helpers
def fr_to_mont(x: int) -> int:
    return x * FR_R % R_MOD


def fr_from_mont(x: int) -> int:
    return x * FR_RINV % R_MOD


def fq_to_mont(x: int) -> int:
    return x * FQ_R % Q_MOD


def fq_from_mont(x: int) -> int:
    return x * FQ_RINV % Q_MOD


def to_limbs(x: int, n: int):
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def from_limbs(limbs) -> int:
    return sum(int(l) << (64 * i) for i, l in enumerate(limbs))


# ----------------------------------------------------------------------------
# radix-2 domain (algebra/poly/src/domain/radix2/mod.rs:51-82)
def k_adicity(k: int, n: int) -> int:
    r = 0
    while n > 1:
        if n % k == 0:
            r += 1
            n //= k
        else:
            return r
    return r


def fr_root_of_unity(n: int) -> int:
    """get_root_of_unity(n), large-subgroup branch (ff/src/fields/mod.rs:339-367)."""
    q = FR_SMALL_SUBGROUP_BASE
    q_adicity = k_adicity(q, n)
    q_part = q**q_adicity
    two_adicity = k_adicity(2, n)
    two_part = 1 << two_adicity
    if n != two_part * q_part or two_adicity > FR_TWO_ADICITY or q_adicity > FR_SMALL_SUBGROUP_BASE_ADICITY:
        raise ValueError("no root of unity of that order")
    omega = FR_LARGE_SUBGROUP_ROOT
    for _ in range(q_adicity, FR_SMALL_SUBGROUP_BASE_ADICITY):
        omega = pow(omega, q, R_MOD)
    for _ in range(two_adicity, FR_TWO_ADICITY):
        omega = omega * omega % R_MOD
    return omega


class Domain:
    def __init__(self, num_coeffs: int):
        size = 1
        while size < num_coeffs:
            size *= 2
        self.size = size
        self.log_size = size.bit_length() - 1
        self.group_gen = fr_root_of_unity(size)
        self.group_gen_inv = pow(self.group_gen, -1, R_MOD)
        self.size_inv = pow(size, -1, R_MOD)
        self.generator_inv = pow(FR_GENERATOR, -1, R_MOD)

    def vanishing_at(self, tau: int) -> int:
        return (pow(tau, self.size, R_MOD) - 1) % R_MOD


def _fast_dft(a, w, p):
    """Recursive radix-2 DFT on Python ints: out[i] = sum a[j] w^(ij)."""
    n = len(a)
    if n == 1:
        return list(a)
    w2 = w * w % p
    ev = _fast_dft(a[0::2], w2, p)
    od = _fast_dft(a[1::2], w2, p)
    out = [0] * n
    t = 1
    h = n // 2
    for i in range(h):
        x = t * od[i] % p
        out[i] = (ev[i] + x) % p
        out[i + h] = (ev[i] - x) % p
        t = t * w % p
    return out


def ntt(values, inverse=False, coset=False, slow=False):
    """The four reference transforms on canonical ints, natural order in and out.

    fft:        out[i] = p(w^i)                                 (radix2/mod.rs:99-103)
    coset fft:  out[i] = p(g w^i)                               (domain/mod.rs:139-142)
    ifft:       inverse of fft                                  (radix2/fft.rs:26-29)
    coset ifft: ifft then coefficient i times g^-i              (radix2/fft.rs:31-35)
    Input shorter than the domain is zero padded (radix2/mod.rs:100-101).
    """
    d = Domain(max(len(values), 1))
    a = [v % R_MOD for v in values] + [0] * (d.size - len(values))
    n = d.size
    if not inverse:
        if coset:
            gp = 1
            for i in range(n):
                a[i] = a[i] * gp % R_MOD
                gp = gp * FR_GENERATOR % R_MOD
        if slow:
            return [sum(a[j] * pow(d.group_gen, i * j, R_MOD) for j in range(n)) % R_MOD for i in range(n)]
        return _fast_dft(a, d.group_gen, R_MOD)
    if slow:
        out = [sum(a[j] * pow(d.group_gen_inv, i * j, R_MOD) for j in range(n)) % R_MOD for i in range(n)]
    else:
        out = _fast_dft(a, d.group_gen_inv, R_MOD)
    out = [x * d.size_inv % R_MOD for x in out]
    if coset:
        gp = 1
        for i in range(n):
            out[i] = out[i] * gp % R_MOD
            gp = gp * d.generator_inv % R_MOD
    return out


# ----------------------------------------------------------------------------
# Fq2 = Fq[u]/(u^2 + 5)
def fq2_add(a, b):
    return ((a[0] + b[0]) % Q_MOD, (a[1] + b[1]) % Q_MOD)


def fq2_sub(a, b):
    return ((a[0] - b[0]) % Q_MOD, (a[1] - b[1]) % Q_MOD)


def fq2_mul(a, b):
    return ((a[0] * b[0] + FQ2_NONRESIDUE * a[1] * b[1]) % Q_MOD, (a[0] * b[1] + a[1] * b[0]) % Q_MOD)


def fq2_inv(a):
    # 1/(c0 + c1 u) = (c0 - c1 u)/(c0^2 - beta c1^2)
    norm = (a[0] * a[0] - FQ2_NONRESIDUE * a[1] * a[1]) % Q_MOD
    ni = pow(norm, -1, Q_MOD)
    return (a[0] * ni % Q_MOD, (-a[1]) * ni % Q_MOD)


class _F1:
    zero = 0
    one = 1
    add = staticmethod(lambda a, b: (a + b) % Q_MOD)
    sub = staticmethod(lambda a, b: (a - b) % Q_MOD)
    mul = staticmethod(lambda a, b: a * b % Q_MOD)
    inv = staticmethod(lambda a: pow(a, -1, Q_MOD))
    neg = staticmethod(lambda a: (-a) % Q_MOD)
    b = G1_B


class _F2:
    zero = (0, 0)
    one = (1, 0)
    add = staticmethod(fq2_add)
    sub = staticmethod(fq2_sub)
    mul = staticmethod(fq2_mul)
    inv = staticmethod(fq2_inv)
    neg = staticmethod(lambda a: ((-a[0]) % Q_MOD, (-a[1]) % Q_MOD))
    b = G2_B


# affine points: None == infinity, else (x, y)
def _on_curve(F, P):
    if P is None:
        return True
    x, y = P
    return F.mul(y, y) == F.add(F.mul(F.mul(x, x), x), F.b)


def _add(F, P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if y1 != y2 or y1 == F.zero:
            return None
        xx = F.mul(x1, x1)
        lam = F.mul(F.add(F.add(xx, xx), xx), F.inv(F.add(y1, y1)))
    else:
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
    y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
    return (x3, y3)


def _neg(F, P):
    return None if P is None else (P[0], F.neg(P[1]))


def _mul(F, P, k: int):
    k %= R_MOD
    acc = None
    while k:
        if k & 1:
            acc = _add(F, acc, P)
        P = _add(F, P, P)
        k >>= 1
    return acc


def g1_on_curve(P):
    return _on_curve(_F1, P)


def g1_add(P, Q):
    return _add(_F1, P, Q)


def g1_neg(P):
    return _neg(_F1, P)


def g1_mul(P, k):
    return _mul(_F1, P, k)


def g2_on_curve(P):
    return _on_curve(_F2, P)


def g2_add(P, Q):
    return _add(_F2, P, Q)


def g2_neg(P):
    return _neg(_F2, P)


def g2_mul(P, k):
    return _mul(_F2, P, k)


def g1_msm_naive(bases, scalars):
    """sum s_i P_i on affine points (algebra/test-templates/src/msm.rs:6-14)."""
    acc = None
    for P, s in zip(bases, scalars):
        acc = g1_add(acc, g1_mul(P, s))
    return acc


def g2_msm_naive(bases, scalars):
    acc = None
    for P, s in zip(bases, scalars):
        acc = g2_add(acc, g2_mul(P, s))
    return acc


# ----------------------------------------------------------------------------
# reference window heuristic (ec/src/msm/variable_base.rs:21-25, msm/mod.rs:10-13, utils/src/lib.rs:65-73)
def ark_log2(x: int) -> int:
    if x == 0:
        return 0
    if x & (x - 1) == 0:
        return x.bit_length() - 1
    return x.bit_length()


def ref_window(size: int) -> int:
    return 3 if size < 32 else ark_log2(size) * 69 // 100 + 2


def ref_msm_adds(n: int) -> int:
    """The metric normaliser of SURVEY.md section 8(d): N*W + 2(2^c-1)W + 253."""
    c = ref_window(n)
    w = (253 + c - 1) // c
    return n * w + 2 * ((1 << c) - 1) * w + 253


# ----------------------------------------------------------------------------
# ark-serialize canonical encodings, restated on Python integers (test infrastructure): the independent model the
# product's czk_*_serialize / czk_*_deserialize are checked against.
#   algebra/ff/src/fields/macros.rs:1-87 (Fp), fields/models/quadratic_extension.rs:600-647 (Fq2: c0 | c1),
#   algebra/serialize/src/flags.rs (SWFlags: bit 7 = y > -y, bit 6 = infinity),
#   algebra/ec/src/models/short_weierstrass_jacobian.rs:108-118, 792-895 (GroupAffine),
#   orderings: macros.rs:507-512 (Fp by canonical integer), quadratic_extension.rs:410-419 (Fq2 by (c1, c0)).
SER_POSITIVE_Y, SER_INFINITY = 1 << 7, 1 << 6


def ser_fr(x: int) -> bytes:
    return (x % R_MOD).to_bytes(32, "little")


def deser_fr(b: bytes):
    v = int.from_bytes(b, "little")
    return v if v < R_MOD else None  # also rejects a set top bit (EmptyFlags::from_u8)


def _fq_bytes(x: int) -> bytearray:
    return bytearray((x % Q_MOD).to_bytes(48, "little"))


def _coord_bytes(g: int, x) -> bytearray:
    return _fq_bytes(x) if g == 1 else _fq_bytes(x[0]) + _fq_bytes(x[1])


def _coord_key(g: int, y):
    return y if g == 1 else (y[1], y[0])


def _F(g: int):
    return _F1 if g == 1 else _F2


def ser_point(g: int, P, compressed=True) -> bytes:
    """g: 1 = G1, 2 = G2; P: None (infinity) or affine (x, y)."""
    F = _F(g)
    if compressed:
        if P is None:
            out = _coord_bytes(g, F.zero)
            out[-1] |= SER_INFINITY
            return bytes(out)
        out = _coord_bytes(g, P[0])
        if _coord_key(g, P[1]) > _coord_key(g, F.neg(P[1])):
            out[-1] |= SER_POSITIVE_Y
        return bytes(out)
    x, y = (F.zero, F.one) if P is None else P
    out = _coord_bytes(g, x) + _coord_bytes(g, y)
    if P is None:
        out[-1] |= SER_INFINITY
    return bytes(out)


def fq_sqrt(a: int):
    """Tonelli-Shanks; None for a non-residue."""
    a %= Q_MOD
    if a == 0:
        return 0
    if pow(a, (Q_MOD - 1) // 2, Q_MOD) != 1:
        return None
    s, t = 0, Q_MOD - 1
    while t % 2 == 0:
        s, t = s + 1, t // 2
    g = next(k for k in range(2, 100) if pow(k, (Q_MOD - 1) // 2, Q_MOD) != 1)
    z, x, b, m = pow(g, t, Q_MOD), pow(a, (t + 1) // 2, Q_MOD), pow(a, t, Q_MOD), s
    while b != 1:
        k, b2 = 0, b
        while b2 != 1:
            b2, k = b2 * b2 % Q_MOD, k + 1
        w = pow(z, 1 << (m - k - 1), Q_MOD)
        z, b, x, m = w * w % Q_MOD, b * w * w % Q_MOD, x * w % Q_MOD, k
    return x


def fq2_sqrt(a):
    """quadratic_extension.rs:360-399 (the complex method), including its None for c1 = 0 with c0 a non-residue."""
    c0, c1 = a[0] % Q_MOD, a[1] % Q_MOD
    if c1 == 0:
        r = fq_sqrt(c0)
        return None if r is None else (r, 0)
    alpha = fq_sqrt((c0 * c0 - FQ2_NONRESIDUE * c1 * c1) % Q_MOD)
    if alpha is None:
        return None
    two_inv = pow(2, -1, Q_MOD)
    delta = (alpha + c0) * two_inv % Q_MOD
    if fq_sqrt(delta) is None:
        delta = (delta - alpha) % Q_MOD
    r0 = fq_sqrt(delta)
    if not r0:
        return None
    cand = (r0, c1 * two_inv % Q_MOD * pow(r0, -1, Q_MOD) % Q_MOD)
    return cand if fq2_mul(cand, cand) == (c0, c1) else None


def deser_point(g: int, b: bytes, compressed=True, check_subgroup=True):
    """Returns ('ok', P) or ('err', reason) the way GroupAffine::deserialize accepts / rejects."""
    F = _F(g)
    n = 48 * g
    buf = bytearray(b)
    pos, inf = bool(buf[-1] & SER_POSITIVE_Y), bool(buf[-1] & SER_INFINITY)
    if pos and inf:
        return "err", "flags"
    buf[-1] &= 0xFF ^ (SER_POSITIVE_Y | SER_INFINITY)

    def coord(chunk):
        vals = [int.from_bytes(chunk[i:i + 48], "little") for i in range(0, len(chunk), 48)]
        if any(v >= Q_MOD for v in vals):
            return None
        return vals[0] if g == 1 else (vals[0], vals[1])

    x = coord(buf[:n])
    if x is None:
        return "err", "field"
    if compressed:
        if inf:
            return "ok", None
        rhs = F.add(F.mul(F.mul(x, x), x), F.b)
        y = fq_sqrt(rhs) if g == 1 else fq2_sqrt(rhs)
        if y is None:
            return "err", "curve"
        negy = F.neg(y)
        if not ((_coord_key(g, y) < _coord_key(g, negy)) ^ pos):
            y = negy
    else:
        if pos:
            return "err", "flags"
        y = coord(buf[n:])
        if y is None:
            return "err", "field"
        if inf:
            return "ok", None
        if not _on_curve(F, (x, y)):
            return "err", "curve"
    if check_subgroup and _mul_raw(F, (x, y), R_MOD) is not None:
        return "err", "subgroup"
    return "ok", (x, y)


def _mul_raw(F, P, k: int):
    """k * P without reducing k modulo the group order (the subgroup test multiplies by r itself)."""
    acc = None
    while k:
        if k & 1:
            acc = _add(F, acc, P)
        P = _add(F, P, P)
        k >>= 1
    return acc


# ----------------------------------------------------------------------------
# BLS12-377 ate pairing on Python integers (test infrastructure): the model the product's host-side verifier
# (czk_pairing_product_is_one / czk_groth16_verify) is checked against.  The reference computes the same bilinear map with a
# projective Miller loop and a cyclotomic final exponentiation (algebra/ec/src/models/bls12/mod.rs:59-200,
# curves/bls12_377/src/curves/mod.rs: x = 0x8508c00000000001, D-type twist over Fq2 with xi = u); here it is the textbook form:
# untwist Q into E(Fq12), affine Miller loop over t - 1 = x, and f^((q^12 - 1) / r) by plain square-and-multiply.  Same
# function up to a fixed non-zero power (gcd-free exponents differ), so pairing-PRODUCT checks - all a verifier needs - agree.
# Tower: Fq2 = Fq[u]/(u^2 + 5), Fq6 = Fq2[v]/(v^3 - u), Fq12 = Fq6[w]/(w^2 - v)   (fields/fq6.rs, fq12.rs).
BLS_X = 0x8508C00000000001
FQ2_ZERO, FQ2_ONE = (0, 0), (1, 0)


def fq2_mul_xi(a):  # times xi = u
    return (FQ2_NONRESIDUE * a[1] % Q_MOD, a[0])


def fq6_add(a, b):
    return tuple(fq2_add(x, y) for x, y in zip(a, b))


def fq6_sub(a, b):
    return tuple(fq2_sub(x, y) for x, y in zip(a, b))


def fq6_mul(a, b):
    a0, a1, a2 = a
    b0, b1, b2 = b
    c0 = fq2_add(fq2_mul(a0, b0), fq2_mul_xi(fq2_add(fq2_mul(a1, b2), fq2_mul(a2, b1))))
    c1 = fq2_add(fq2_add(fq2_mul(a0, b1), fq2_mul(a1, b0)), fq2_mul_xi(fq2_mul(a2, b2)))
    c2 = fq2_add(fq2_add(fq2_mul(a0, b2), fq2_mul(a1, b1)), fq2_mul(a2, b0))
    return (c0, c1, c2)


def fq6_mul_v(a):
    return (fq2_mul_xi(a[2]), a[0], a[1])


def fq6_inv(a):
    a0, a1, a2 = a
    t0 = fq2_sub(fq2_mul(a0, a0), fq2_mul_xi(fq2_mul(a1, a2)))
    t1 = fq2_sub(fq2_mul_xi(fq2_mul(a2, a2)), fq2_mul(a0, a1))
    t2 = fq2_sub(fq2_mul(a1, a1), fq2_mul(a0, a2))
    d = fq2_add(fq2_mul(a0, t0), fq2_mul_xi(fq2_add(fq2_mul(a2, t1), fq2_mul(a1, t2))))
    di = fq2_inv(d)
    return (fq2_mul(t0, di), fq2_mul(t1, di), fq2_mul(t2, di))


FQ6_ZERO = (FQ2_ZERO, FQ2_ZERO, FQ2_ZERO)
FQ6_ONE = (FQ2_ONE, FQ2_ZERO, FQ2_ZERO)
FQ12_ONE = (FQ6_ONE, FQ6_ZERO)


def fq12_add(a, b):
    return (fq6_add(a[0], b[0]), fq6_add(a[1], b[1]))


def fq12_sub(a, b):
    return (fq6_sub(a[0], b[0]), fq6_sub(a[1], b[1]))


def fq12_mul(a, b):
    t0, t1 = fq6_mul(a[0], b[0]), fq6_mul(a[1], b[1])
    c1 = fq6_sub(fq6_sub(fq6_mul(fq6_add(a[0], a[1]), fq6_add(b[0], b[1])), t0), t1)
    return (fq6_add(t0, fq6_mul_v(t1)), c1)


def fq12_inv(a):
    d = fq6_sub(fq6_mul(a[0], a[0]), fq6_mul_v(fq6_mul(a[1], a[1])))
    di = fq6_inv(d)
    return (fq6_mul(a[0], di), fq6_sub(FQ6_ZERO, fq6_mul(a[1], di)))


def fq12_pow(a, e: int):
    res, started = FQ12_ONE, False
    for i in range(e.bit_length() - 1, -1, -1):
        if started:
            res = fq12_mul(res, res)
        if (e >> i) & 1:
            res = fq12_mul(res, a)
            started = True
    return res


def fq12_from_fq(x: int):
    return (((x % Q_MOD, 0), FQ2_ZERO, FQ2_ZERO), FQ6_ZERO)


def g2_untwist(Q):
    """(x', y') on the twist -> (x' w^2, y' w^3) on E(Fq12): w^2 = v, w^3 = v w."""
    x, y = Q
    return ((FQ2_ZERO, x, FQ2_ZERO), FQ6_ZERO), (FQ6_ZERO, (FQ2_ZERO, y, FQ2_ZERO))


def miller_loop(P, Q):
    """f_{x, psi(Q)}(P) for affine P in G1, Q in G2 (neither infinity), without the vertical lines (they lie in Fq6 and
    die in the final exponentiation)."""
    xp, yp = fq12_from_fq(P[0]), fq12_from_fq(P[1])
    xq, yq = g2_untwist(Q)
    xt, yt = xq, yq
    f = FQ12_ONE
    three, two = fq12_from_fq(3), fq12_from_fq(2)

    def line(lam, x0, y0):  # (y_P - y0) - lam (x_P - x0)
        return fq12_sub(fq12_sub(yp, y0), fq12_mul(lam, fq12_sub(xp, x0)))

    for i in range(BLS_X.bit_length() - 2, -1, -1):
        lam = fq12_mul(fq12_mul(three, fq12_mul(xt, xt)), fq12_inv(fq12_mul(two, yt)))
        f = fq12_mul(fq12_mul(f, f), line(lam, xt, yt))
        x3 = fq12_sub(fq12_sub(fq12_mul(lam, lam), xt), xt)
        yt = fq12_sub(fq12_mul(lam, fq12_sub(xt, x3)), yt)
        xt = x3
        if (BLS_X >> i) & 1:
            lam = fq12_mul(fq12_sub(yq, yt), fq12_inv(fq12_sub(xq, xt)))
            f = fq12_mul(f, line(lam, xt, yt))
            x3 = fq12_sub(fq12_sub(fq12_mul(lam, lam), xt), xq)
            yt = fq12_sub(fq12_mul(lam, fq12_sub(xt, x3)), yt)
            xt = x3
    return f


FINAL_EXP = (Q_MOD ** 12 - 1) // R_MOD


def pairing(P, Q):
    """e(P, Q) in Fq12 (1 if either argument is infinity)."""
    if P is None or Q is None:
        return FQ12_ONE
    return fq12_pow(miller_loop(P, Q), FINAL_EXP)


def pairing_product_is_one(pairs) -> bool:
    """prod e(P_i, Q_i) == 1 with ONE final exponentiation (what a Groth16 verifier evaluates)."""
    f = FQ12_ONE
    for P, Q in pairs:
        if P is not None and Q is not None:
            f = fq12_mul(f, miller_loop(P, Q))
    return fq12_pow(f, FINAL_EXP) == FQ12_ONE


def groth16_verify(vk, proof, public_inputs) -> bool:
    """groth16/src/verifier.rs: e(A, B) == e(alpha, beta) e(sum x_i gamma_abc_i, gamma) e(C, delta), as one product.
    vk: dict(alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1 list) of affine int points; public_inputs without the one."""
    A, B, C = proof
    acc = vk["gamma_abc_g1"][0]
    for x, pt in zip(public_inputs, vk["gamma_abc_g1"][1:]):
        acc = g1_add(acc, g1_mul(pt, x))
    return pairing_product_is_one([(A, B), (g1_neg(vk["alpha_g1"]), vk["beta_g2"]), (g1_neg(acc), vk["gamma_g2"]),
                                   (g1_neg(C), vk["delta_g2"])])
