// Host build of the DEVICE math headers (csrc/fp.cuh, csrc/ec.cuh) with the PTX carry flag
// emulated (csrc/carry.cuh).  Test infrastructure: lets the CPU-only test tier check the exact
// limb schedule the GPU runs against the oracle.  Not part of the product library.
#include <cstring>
#include "../../collaborative-zksnark_b200/csrc/ec.cuh"
#include "../../collaborative-zksnark_b200/csrc/msm_digits.cuh"
#include "../../collaborative-zksnark_b200/csrc/fq13.cuh"
#include "../../collaborative-zksnark_b200/csrc/fq_inverse.cuh"

using namespace czk;
#define EXPORT extern "C" __attribute__((visibility("default")))

template <class F>
static F ld(const uint64_t* p) {
    F r;
    std::memcpy(r.l, p, sizeof r.l);
    return r;
}
template <class F>
static void st(uint64_t* p, const F& v) {
    std::memcpy(p, v.l, sizeof v.l);
}
static Fq2 ld2(const uint64_t* p) { return Fq2{ld<Fq>(p), ld<Fq>(p + 6)}; }
static void st2(uint64_t* p, const Fq2& v) {
    st(p, v.c0);
    st(p + 6, v.c1);
}

EXPORT void emu_fr_mul(uint64_t* r, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 4 * i, Fr::mul(ld<Fr>(a + 4 * i), ld<Fr>(b + 4 * i)));
}
EXPORT void emu_fr_add(uint64_t* r, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 4 * i, Fr::add(ld<Fr>(a + 4 * i), ld<Fr>(b + 4 * i)));
}
EXPORT void emu_fr_sub(uint64_t* r, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 4 * i, Fr::sub(ld<Fr>(a + 4 * i), ld<Fr>(b + 4 * i)));
}
EXPORT void emu_fr_neg(uint64_t* r, const uint64_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 4 * i, Fr::neg(ld<Fr>(a + 4 * i)));
}
EXPORT void emu_fr_from_mont(uint64_t* r, const uint64_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 4 * i, Fr::from_mont(ld<Fr>(a + 4 * i)));
}
EXPORT void emu_fr_inv(uint64_t* r, const uint64_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 4 * i, Fr::inv_fermat(ld<Fr>(a + 4 * i)));
}
EXPORT void emu_fq_mul(uint64_t* r, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 6 * i, Fq::mul(ld<Fq>(a + 6 * i), ld<Fq>(b + 6 * i)));
}
EXPORT void emu_fq_add(uint64_t* r, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 6 * i, Fq::add(ld<Fq>(a + 6 * i), ld<Fq>(b + 6 * i)));
}
EXPORT void emu_fq_sub(uint64_t* r, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 6 * i, Fq::sub(ld<Fq>(a + 6 * i), ld<Fq>(b + 6 * i)));
}
EXPORT void emu_fq_inv(uint64_t* r, const uint64_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 6 * i, Fq::inv_fermat(ld<Fq>(a + 6 * i)));
}
EXPORT void emu_fq_mul2(uint64_t* r1, uint64_t* r2, const uint64_t* a, const uint64_t* b, const uint64_t* c, const uint64_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) {
        Fq x, y;
        Fq::mul2(ld<Fq>(a + 6 * i), ld<Fq>(b + 6 * i), ld<Fq>(c + 6 * i), ld<Fq>(d + 6 * i), x, y);
        st(r1 + 6 * i, x);
        st(r2 + 6 * i, y);
    }
}
// the batched-step binary GCD inversion (csrc/fq_inverse.cuh)
EXPORT void emu_fq_inv_bingcd(uint64_t* r, const uint64_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 6 * i, BinGcd<FqParams>::inverse(ld<Fq>(a + 6 * i)));
}
EXPORT void emu_fr_inv_bingcd(uint64_t* r, const uint64_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 4 * i, BinGcd<FrParams>::inverse(ld<Fr>(a + 4 * i)));
}
EXPORT void emu_fq2_mul(uint64_t* r, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) st2(r + 12 * i, Fq2::mul(ld2(a + 12 * i), ld2(b + 12 * i)));
}
EXPORT void emu_fq2_sqr(uint64_t* r, const uint64_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) st2(r + 12 * i, Fq2::sqr(ld2(a + 12 * i)));
}
EXPORT void emu_fq2_inv(uint64_t* r, const uint64_t* a, size_t n) {
    for (size_t i = 0; i < n; i++) st2(r + 12 * i, Fq2::inv_fermat(ld2(a + 12 * i)));
}

// sum of affine points with signs, accumulated exactly like a device bucket; result affine
// (via the device Fermat inversion).  sign[i] != 0 negates point i.  returns 1 if infinity.
EXPORT int emu_g1_sum(uint64_t* out_xy, const uint64_t* xy, const uint8_t* sign, size_t n, int tree) {
    XYZZ<Fq> acc = XYZZ<Fq>::infinity();
    if (!tree) {
        for (size_t i = 0; i < n; i++) {
            Fq x = ld<Fq>(xy + 12 * i), y = ld<Fq>(xy + 12 * i + 6);
            if (sign && sign[i]) y = Fq::neg(y);
            acc.add_affine(x, y);
        }
    } else {  // exercise the XYZZ+XYZZ path: pairwise then fold
        for (size_t i = 0; i < n; i += 2) {
            XYZZ<Fq> t = XYZZ<Fq>::infinity();
            for (size_t k = i; k < i + 2 && k < n; k++) {
                Fq x = ld<Fq>(xy + 12 * k), y = ld<Fq>(xy + 12 * k + 6);
                if (sign && sign[k]) y = Fq::neg(y);
                t.add_affine(x, y);
            }
            acc.add(t);
        }
    }
    if (acc.is_inf()) return 1;
    Fq zi = Fq::inv_fermat(acc.zz), zzzi = Fq::inv_fermat(acc.zzz);
    st(out_xy, Fq::mul(acc.x, zi));
    st(out_xy + 6, Fq::mul(acc.y, zzzi));
    return 0;
}
EXPORT int emu_g2_sum(uint64_t* out_xy, const uint64_t* xy, const uint8_t* sign, size_t n, int tree) {
    XYZZ<Fq2> acc = XYZZ<Fq2>::infinity();
    for (size_t i = 0; i < n; i += 2) {
        XYZZ<Fq2> t = XYZZ<Fq2>::infinity();
        for (size_t k = i; k < i + 2 && k < n; k++) {
            Fq2 x = ld2(xy + 24 * k), y = ld2(xy + 24 * k + 12);
            if (sign && sign[k]) y = Fq2::neg(y);
            if (tree) t.add_affine(x, y);
            else acc.add_affine(x, y);
        }
        if (tree) acc.add(t);
    }
    if (acc.is_inf()) return 1;
    Fq2 zi = Fq2::inv_fermat(acc.zz), zzzi = Fq2::inv_fermat(acc.zzz);
    st2(out_xy, Fq2::mul(acc.x, zi));
    st2(out_xy + 12, Fq2::mul(acc.y, zzzi));
    return 0;
}
// signed-digit decomposition used by the MSM kernels (csrc/msm_digits.cuh)
EXPORT void emu_signed_digits(int32_t* out, const uint64_t* scalar_canonical, unsigned c, unsigned nwin) {
    uint32_t s[8];
    std::memcpy(s, scalar_canonical, 32);
    signed_digits(s, c, nwin, out);
}

// 13 x 29-bit digit field (csrc/fq13.cuh): op on values given / returned in the reference (12-limb, R = 2^384) form
EXPORT void emu_fq13_binop(uint64_t* r, const uint64_t* a, const uint64_t* b, size_t n, int op) {
    for (size_t i = 0; i < n; i++) {
        Fq13 x = Fq13::from_std(ld<Fq>(a + 6 * i)), y = Fq13::from_std(ld<Fq>(b + 6 * i));
        Fq13 z = op == 0 ? Fq13::mul(x, y) : op == 1 ? Fq13::add(x, y) : op == 2 ? Fq13::sub(x, y) : Fq13::neg(x);
        st(r + 6 * i, z.to_std());
    }
}
EXPORT int emu_g1_sum13(uint64_t* out_xy, const uint64_t* xy, const uint8_t* sign, size_t n) {
    XYZZ<Fq13> acc = XYZZ<Fq13>::infinity();
    for (size_t i = 0; i < n; i++) {
        Fq13 x = Fq13::from_std(ld<Fq>(xy + 12 * i)), y = Fq13::from_std(ld<Fq>(xy + 12 * i + 6));
        if (sign && sign[i]) y = Fq13::neg(y);
        acc.add_affine(x, y);
    }
    if (acc.is_inf()) return 1;
    XYZZ<Fq> s{acc.x.to_std(), acc.y.to_std(), acc.zz.to_std(), acc.zzz.to_std()};
    Fq zi = Fq::inv_fermat(s.zz), zzzi = Fq::inv_fermat(s.zzz);
    st(out_xy, Fq::mul(s.x, zi));
    st(out_xy + 6, Fq::mul(s.y, zzzi));
    return 0;
}
// The two-lane Fq2 product of the G2 bucket kernel (csrc/msm_batched.cu): lane 0 owns c0, lane 1 owns c1; each lane
// computes ONE lazily reduced sum of two products (Fp::mul_sum2).  Here both lanes are evaluated one after the other.
//   c0 = a0 b0 + (-5 a1) b1      c1 = a1 b0 + a0 b1
EXPORT void emu_fq2_mul_two_lanes(uint64_t* r, const uint64_t* a, const uint64_t* b, size_t n) {
    for (size_t i = 0; i < n; i++) {
        Fq2 x = ld2(a + 12 * i), y = ld2(b + 12 * i);
        uint32_t m5[12];
        Fq::neg_times5_unreduced(x.c1, m5);
        Fq2 out;
        out.c0 = Fq::mul_sum2(x.c0, y.c0, m5, y.c1);
        out.c1 = Fq::mul_sum2(x.c1, y.c0, x.c0.l, y.c1);
        st2(r + 12 * i, out);
    }
}
EXPORT void emu_fr_mul_sum2(uint64_t* r, const uint64_t* a, const uint64_t* b, const uint64_t* c, const uint64_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) st(r + 4 * i, Fr::mul_sum2(ld<Fr>(a + 4 * i), ld<Fr>(b + 4 * i), ld<Fr>(c + 4 * i).l, ld<Fr>(d + 4 * i)));
}
