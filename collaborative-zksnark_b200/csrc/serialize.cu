// ark-serialize canonical wire format of Fr, G1Affine, G2Affine and the Groth16 proof:
// what a GPU party must read and write to talk to a stock mpc-net party, and the bytes `Proof::serialize` produces.
// Host code (no kernels): O(1) elements per message.
//
//   algebra/ff/src/fields/macros.rs:1-87              Fp: canonical integer (out of Montgomery form), little-endian bytes,
//                                                     flags OR-ed into the top bits of the last byte
//   algebra/ff/src/fields/models/quadratic_extension.rs:600-647   Fq2: c0 then c1, the flags on c1
//   algebra/serialize/src/flags.rs                    SWFlags: bit 7 = "y is the greater of (y, -y)", bit 6 = infinity
//   algebra/ec/src/models/short_weierstrass_jacobian.rs:792-895   GroupAffine: compressed = x with flags (infinity: x = 0);
//                                                     uncompressed = x, then y with the infinity flag; deserialisation
//                                                     recovers y from x (get_point_from_x, :108-118) and checks the subgroup
//   ordering: Fp by canonical integer (macros.rs:507-512), Fq2 by (c1, c0) (quadratic_extension.rs:410-419)
//   groth16/src/data_structures.rs                    Proof { a: G1Affine, b: G2Affine, c: G1Affine }, compressed: 48 + 96 + 48 bytes
#include "../../include/czk_groth16.h"
#include "ctx.hpp"

namespace {

constexpr uint8_t FLAG_POSITIVE_Y = 1u << 7, FLAG_INFINITY = 1u << 6;

template <class HF>
void fp_to_bytes(const HF& mont, uint8_t* out) {
    HF c = mont.from_mont();
    for (int i = 0; i < HF::N; i++)
        for (int b = 0; b < 8; b++) out[8 * i + b] = (uint8_t)(c.l[i] >> (8 * b));
}
// canonical little-endian bytes -> Montgomery form; false if the integer is >= p
template <class HF>
bool fp_from_bytes(const uint8_t* in, HF* out) {
    HF c;
    for (int i = 0; i < HF::N; i++) {
        uint64_t v = 0;
        for (int b = 0; b < 8; b++) v |= (uint64_t)in[8 * i + b] << (8 * b);
        c.l[i] = v;
    }
    if (HF::geq_mod(c.l)) return false;
    *out = c.to_mont();
    return true;
}
// compare the canonical integers: < 0, 0, > 0
template <class HF>
int fp_cmp(const HF& a, const HF& b) {
    HF x = a.from_mont(), y = b.from_mont();
    for (int i = HF::N - 1; i >= 0; i--) {
        if (x.l[i] != y.l[i]) return x.l[i] < y.l[i] ? -1 : 1;
    }
    return 0;
}
int fq2_cmp(const HFq2& a, const HFq2& b) {
    int c = fp_cmp(a.c1, b.c1);
    return c ? c : fp_cmp(a.c0, b.c0);
}

// ---- square roots (only which of the two roots is returned differs between algorithms; the caller picks by sign)
// Tonelli-Shanks in Fq: q - 1 = 2^46 t
bool fq_sqrt(const HFq& a, HFq* out) {
    if (a.is_zero()) {
        *out = a;
        return true;
    }
    constexpr int S = 46;
    // t = (q - 1) >> 46 ; (t + 1) / 2 ; (q - 1) / 2
    uint64_t qm1[6], t[6], th[6], half[6];
    for (int i = 0; i < 6; i++) qm1[i] = FqParams::MOD64[i];
    qm1[0] -= 1;
    for (int i = 0; i < 6; i++) {
        half[i] = (qm1[i] >> 1) | (i + 1 < 6 ? qm1[i + 1] << 63 : 0);
        int w = S / 64, sh = S % 64;  // S < 64
        (void)w;
        t[i] = (qm1[i] >> sh) | (i + 1 < 6 ? qm1[i + 1] << (64 - sh) : 0);
    }
    if (HFq::pow(a, half, 6) != HFq::one()) return false;  // Euler: not a square
    // (t + 1) / 2 with t odd: t >> 1 plus one
    for (int i = 0; i < 6; i++) th[i] = (t[i] >> 1) | (i + 1 < 6 ? t[i + 1] << 63 : 0);
    {
        int i = 0;
        while (i < 6 && ++th[i] == 0) i++;
    }
    // a quadratic non-residue: the smallest small integer failing Euler's criterion
    HFq g = HFq::zero();
    for (uint64_t k = 2; k < 64; k++) {
        HFq c = HFq::from_u64(k);
        if (HFq::pow(c, half, 6) != HFq::one()) {
            g = c;
            break;
        }
    }
    HFq z = HFq::pow(g, t, 6);  // generator of the 2-Sylow subgroup
    HFq x = HFq::pow(a, th, 6), b = HFq::pow(a, t, 6);
    int m = S;
    while (b != HFq::one()) {
        int k = 0;
        HFq b2 = b;
        while (b2 != HFq::one()) {
            b2 = HFq::sqr(b2);
            k++;
        }
        HFq w = z;
        for (int i = 0; i < m - k - 1; i++) w = HFq::sqr(w);
        z = HFq::sqr(w);
        b = HFq::mul(b, z);
        x = HFq::mul(x, w);
        m = k;
    }
    *out = x;
    return true;
}
bool fq_is_square(const HFq& a) {
    if (a.is_zero()) return true;
    uint64_t half[6];
    for (int i = 0; i < 6; i++) half[i] = FqParams::MOD64[i];
    half[0] -= 1;
    for (int i = 0; i < 6; i++) half[i] = (half[i] >> 1) | (i + 1 < 6 ? half[i + 1] << 63 : 0);
    return HFq::pow(a, half, 6) == HFq::one();
}
// the complex method (quadratic_extension.rs:360-399)
bool fq2_sqrt(const HFq2& a, HFq2* out) {
    if (a.c1.is_zero()) {
        HFq r;
        // the reference returns None when c0 is not a square in Fq (:363-365), although a root exists in Fq2
        if (!fq_sqrt(a.c0, &r)) return false;
        *out = HFq2{r, HFq::zero()};
        return true;
    }
    // norm = c0^2 + 5 c1^2
    HFq norm = HFq::add(HFq::sqr(a.c0), HFq::mul(HFq::from_u64(5), HFq::sqr(a.c1)));
    HFq alpha;
    if (!fq_sqrt(norm, &alpha)) return false;
    HFq two_inv = HFq::inv(HFq::from_u64(2));
    HFq delta = HFq::mul(HFq::add(alpha, a.c0), two_inv);
    if (!fq_is_square(delta)) delta = HFq::sub(delta, alpha);
    HFq c0;
    if (!fq_sqrt(delta, &c0) || c0.is_zero()) return false;
    HFq2 cand{c0, HFq::mul(HFq::mul(a.c1, two_inv), HFq::inv(c0))};
    if (HFq2::sqr(cand) != a) return false;
    *out = cand;
    return true;
}

// ---- one generic implementation over the base field
struct G1Traits {
    typedef HFq F;
    static constexpr int LIMBS = 6, BYTES = 48;
    static void to_bytes(const F& v, uint8_t* o) { fp_to_bytes(v, o); }
    static bool from_bytes(const uint8_t* i, F* v) { return fp_from_bytes(i, v); }
    static int cmp(const F& a, const F& b) { return fp_cmp(a, b); }
    static bool sqrt(const F& a, F* o) { return fq_sqrt(a, o); }
    static F coeff_b() { return HFq::one(); }
    static F neg(const F& a) { return HFq::neg(a); }
};
struct G2Traits {
    typedef HFq2 F;
    static constexpr int LIMBS = 12, BYTES = 96;
    static void to_bytes(const F& v, uint8_t* o) {
        fp_to_bytes(v.c0, o);
        fp_to_bytes(v.c1, o + 48);
    }
    static bool from_bytes(const uint8_t* i, F* v) { return fp_from_bytes(i, &v->c0) && fp_from_bytes(i + 48, &v->c1); }
    static int cmp(const F& a, const F& b) { return fq2_cmp(a, b); }
    static bool sqrt(const F& a, F* o) { return fq2_sqrt(a, o); }
    static F coeff_b() { return HFq2::from_limbs(CurveConsts::G2_B); }
    static F neg(const F& a) { return HFq2::neg(a); }
};

template <class T>
void point_serialize(const uint64_t* xy, bool inf, bool compressed, uint8_t* out) {
    typedef typename T::F F;
    constexpr int B = T::BYTES;
    if (compressed) {
        if (inf) {  // serialize 0 with the infinity flag
            std::memset(out, 0, B);
            out[B - 1] |= FLAG_INFINITY;
            return;
        }
        F x = F::from_limbs(xy), y = F::from_limbs(xy + T::LIMBS);
        T::to_bytes(x, out);
        if (T::cmp(y, T::neg(y)) > 0) out[B - 1] |= FLAG_POSITIVE_Y;  // SWFlags::from_y_sign(y > -y)
    } else {
        // x, then y with the infinity flag; the affine zero is (0, 1, infinity = true)
        F x = inf ? F::zero() : F::from_limbs(xy), y = inf ? F::one() : F::from_limbs(xy + T::LIMBS);
        T::to_bytes(x, out);
        T::to_bytes(y, out + B);
        if (inf) out[2 * B - 1] |= FLAG_INFINITY;
    }
}
// 0 ok, 1 malformed (flags / non-canonical field element), 2 not on the curve / no such point, 3 outside the subgroup
template <class T>
int point_deserialize(const uint8_t* in, bool compressed, bool check_subgroup, uint64_t* xy, uint8_t* inf) {
    typedef typename T::F F;
    constexpr int B = T::BYTES;
    uint8_t buf[2 * 96];
    std::memcpy(buf, in, compressed ? B : 2 * B);
    uint8_t* last = buf + (compressed ? B : 2 * B) - 1;
    const bool f_pos = (*last & FLAG_POSITIVE_Y) != 0, f_inf = (*last & FLAG_INFINITY) != 0;
    if (f_pos && f_inf) return 1;  // SWFlags::from_u8: (true, true) => None
    *last &= (uint8_t)~(FLAG_POSITIVE_Y | FLAG_INFINITY);
    F x, y;
    if (!T::from_bytes(buf, &x)) return 1;
    if (compressed) {
        if (f_inf) {
            *inf = 1;
            F::zero().to_limbs(xy);
            F::one().to_limbs(xy + T::LIMBS);
            return 0;
        }
        F rhs = F::add(F::mul(F::sqr(x), x), T::coeff_b());
        if (!T::sqrt(rhs, &y)) return 2;
        F negy = T::neg(y);
        // get_point_from_x: if (y < negy) ^ greatest { y } else { negy }
        if (!((T::cmp(y, negy) < 0) ^ f_pos)) y = negy;
    } else {
        if (f_pos) return 1;  // the uncompressed form carries only the infinity flag
        if (!T::from_bytes(buf + B, &y)) return 1;
        if (f_inf) {
            *inf = 1;
            x.to_limbs(xy);
            y.to_limbs(xy + T::LIMBS);
            return 0;
        }
        if (F::sqr(y) != F::add(F::mul(F::sqr(x), x), T::coeff_b())) return 2;
    }
    *inf = 0;
    x.to_limbs(xy);
    y.to_limbs(xy + T::LIMBS);
    if (check_subgroup) {  // is_in_correct_subgroup_assuming_on_curve: r * P == 0
        HPoint<F> p = HPoint<F>::from_affine(x, y);
        if (!HPoint<F>::mul(p, FrParams::MOD64, 4).is_inf()) return 3;
    }
    return 0;
}
int point_error(int rc, const char* what) {
    static const char* msg[] = {"", "malformed encoding (flags or a non-canonical field element)", "not a point of the curve",
                                "point outside the prime-order subgroup"};
    return fail(nullptr, CZK_ERR_ARG, std::string(what) + ": " + msg[rc]);
}

}  // namespace

int czk_fr_serialize(const uint64_t* fr_mont, size_t n, uint8_t* out) {
    if ((!fr_mont || !out) && n) return fail(nullptr, CZK_ERR_ARG, "czk_fr_serialize: null");
    for (size_t i = 0; i < n; i++) fp_to_bytes(HFr::from_limbs(fr_mont + 4 * i), out + 32 * i);
    return CZK_OK;
}
int czk_fr_deserialize(const uint8_t* in, size_t n, uint64_t* fr_mont) {
    if ((!fr_mont || !in) && n) return fail(nullptr, CZK_ERR_ARG, "czk_fr_deserialize: null");
    for (size_t i = 0; i < n; i++) {
        HFr v;
        // EmptyFlags::from_u8 rejects a set top bit; a value >= r is rejected by Fp::read
        if ((in[32 * i + 31] >> 7) || !fp_from_bytes(in + 32 * i, &v))
            return fail(nullptr, CZK_ERR_ARG, "czk_fr_deserialize: not a canonical field element");
        v.to_limbs(fr_mont + 4 * i);
    }
    return CZK_OK;
}
int czk_g1_serialize(const uint64_t* xy, const uint8_t* inf, size_t n, int compressed, uint8_t* out) {
    if ((!xy || !out) && n) return fail(nullptr, CZK_ERR_ARG, "czk_g1_serialize: null");
    const size_t sz = compressed ? 48 : 96;
    for (size_t i = 0; i < n; i++) point_serialize<G1Traits>(xy + 12 * i, inf && inf[i], compressed != 0, out + sz * i);
    return CZK_OK;
}
int czk_g2_serialize(const uint64_t* xy, const uint8_t* inf, size_t n, int compressed, uint8_t* out) {
    if ((!xy || !out) && n) return fail(nullptr, CZK_ERR_ARG, "czk_g2_serialize: null");
    const size_t sz = compressed ? 96 : 192;
    for (size_t i = 0; i < n; i++) point_serialize<G2Traits>(xy + 24 * i, inf && inf[i], compressed != 0, out + sz * i);
    return CZK_OK;
}
int czk_g1_deserialize(const uint8_t* in, size_t n, int compressed, int check_subgroup, uint64_t* xy, uint8_t* inf) {
    if ((!xy || !in || !inf) && n) return fail(nullptr, CZK_ERR_ARG, "czk_g1_deserialize: null");
    const size_t sz = compressed ? 48 : 96;
    for (size_t i = 0; i < n; i++) {
        int rc = point_deserialize<G1Traits>(in + sz * i, compressed != 0, check_subgroup != 0, xy + 12 * i, inf + i);
        if (rc) return point_error(rc, "czk_g1_deserialize");
    }
    return CZK_OK;
}
int czk_g2_deserialize(const uint8_t* in, size_t n, int compressed, int check_subgroup, uint64_t* xy, uint8_t* inf) {
    if ((!xy || !in || !inf) && n) return fail(nullptr, CZK_ERR_ARG, "czk_g2_deserialize: null");
    const size_t sz = compressed ? 96 : 192;
    for (size_t i = 0; i < n; i++) {
        int rc = point_deserialize<G2Traits>(in + sz * i, compressed != 0, check_subgroup != 0, xy + 24 * i, inf + i);
        if (rc) return point_error(rc, "czk_g2_deserialize");
    }
    return CZK_OK;
}
// Proof::serialize: a (G1, 48 bytes) | b (G2, 96 bytes) | c (G1, 48 bytes), compressed
int czk_groth16_proof_serialize(const uint64_t proof[48], const uint8_t proof_inf[3], uint8_t out[192]) {
    if (!proof || !proof_inf || !out) return fail(nullptr, CZK_ERR_ARG, "czk_groth16_proof_serialize: null");
    point_serialize<G1Traits>(proof, proof_inf[0] != 0, true, out);
    point_serialize<G2Traits>(proof + 12, proof_inf[1] != 0, true, out + 48);
    point_serialize<G1Traits>(proof + 36, proof_inf[2] != 0, true, out + 144);
    return CZK_OK;
}
int czk_groth16_proof_deserialize(const uint8_t in[192], uint64_t proof[48], uint8_t proof_inf[3]) {
    if (!proof || !proof_inf || !in) return fail(nullptr, CZK_ERR_ARG, "czk_groth16_proof_deserialize: null");
    int rc;
    if ((rc = point_deserialize<G1Traits>(in, true, true, proof, proof_inf))) return point_error(rc, "proof.a");
    if ((rc = point_deserialize<G2Traits>(in + 48, true, true, proof + 12, proof_inf + 1))) return point_error(rc, "proof.b");
    if ((rc = point_deserialize<G1Traits>(in + 144, true, true, proof + 36, proof_inf + 2))) return point_error(rc, "proof.c");
    return CZK_OK;
}
