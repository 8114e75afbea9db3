// Fq in 13 digits of 29 bits: the representation the bucket-accumulation kernel computes in.
//
// Measured on B200 (tools/mb.py, profiles/): a plain IMAD.WIDE.U32 issues at 64 lanes/clk/SM, but the
// carry-chained form IMAD.WIDE.U32.X that a saturated 32-bit-limb Montgomery product needs runs at ~20.
// With 29-bit digits a 32x32->64 product has 6 spare bits, so a whole column of the schoolbook product
// and of the Montgomery reduction (13 + 13 terms < 2^63) accumulates inside the multiply-add's own 64-bit
// adder: no carry flag anywhere, every multiply is the full-rate instruction, and the 13 column
// accumulators of a row are independent (ILP).  377 = 13 * 29 exactly, so canonical values [0, p) fill the
// digits with no slack and the Montgomery radix is R13 = 2^377.
//
// Values are always canonical (fully reduced, every digit < 2^29): equality and zero tests are digit
// compares, exactly like the 32-bit-limb Fp.  Conversion to and from the reference's in-memory form
// (12 x 32-bit limbs, R = 2^384) is one product by a constant each way (from_std / to_std); bases are
// converted once at upload, bucket sums once at store.
#pragma once
#include "fp.cuh"

namespace czk {

struct Fq13 {
    static constexpr int N = 13;
    static constexpr int B = 29;
    static constexpr uint32_t MASK = Fq13Params::MASK;
    uint32_t d[13];

    CZK_HD static Fq13 zero() {
        Fq13 r;
#pragma unroll
        for (int i = 0; i < 13; i++) r.d[i] = 0;
        return r;
    }
    CZK_HD static Fq13 one() {
        Fq13 r;
#pragma unroll
        for (int i = 0; i < 13; i++) r.d[i] = Fq13Params::one(i);
        return r;
    }
    CZK_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 13; i++) o |= d[i];
        return o == 0;
    }
    CZK_HD bool operator==(const Fq13& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 13; i++) o |= d[i] ^ b.d[i];
        return o == 0;
    }
    CZK_HD bool operator!=(const Fq13& b) const { return !(*this == b); }

    // x in [0, 2p) with digits < 2^29 except the top one (< 2^30)  ->  canonical
    CZK_HD static void cond_sub_p(uint32_t* x) {
        uint32_t s[13];
        int32_t borrow = 0;
#pragma unroll
        for (int k = 0; k < 13; k++) {
            int32_t v = (int32_t)x[k] - (int32_t)Fq13Params::mod(k) - borrow;
            borrow = (v >> 31) & 1;
            s[k] = (uint32_t)v & MASK;
        }
#pragma unroll
        for (int k = 0; k < 13; k++) x[k] = borrow ? x[k] : s[k];
    }
    CZK_HD static Fq13 add(const Fq13& a, const Fq13& b) {
        Fq13 r;
        uint32_t c = 0;
#pragma unroll
        for (int k = 0; k < 12; k++) {
            uint32_t v = a.d[k] + b.d[k] + c;
            r.d[k] = v & MASK;
            c = v >> 29;
        }
        r.d[12] = a.d[12] + b.d[12] + c;  // < 2^30: a + b < 2p
        cond_sub_p(r.d);
        return r;
    }
    CZK_HD static Fq13 sub(const Fq13& a, const Fq13& b) {
        Fq13 r;
        int32_t borrow = 0;
#pragma unroll
        for (int k = 0; k < 13; k++) {
            int32_t v = (int32_t)a.d[k] - (int32_t)b.d[k] - borrow;
            borrow = (v >> 31) & 1;
            r.d[k] = (uint32_t)v & MASK;
        }
        // a < b: the digits now hold a - b + 2^377; add p and drop the 2^377
        uint32_t m = borrow ? 0xffffffffu : 0u, c = 0;
#pragma unroll
        for (int k = 0; k < 13; k++) {
            uint32_t v = r.d[k] + (Fq13Params::mod(k) & m) + c;
            r.d[k] = v & MASK;
            c = v >> 29;
        }
        return r;
    }
    CZK_HD static Fq13 dbl(const Fq13& a) { return add(a, a); }
    CZK_HD static Fq13 neg(const Fq13& a) {
        Fq13 z = zero();
        return sub(z, a);
    }
    // Montgomery product a b / 2^377 mod p.  Column k of the running sum is one 64-bit register; row i adds
    // a_i * b and m_i * p, after which column i is divisible by 2^29 and its quotient moves to column i + 1.
    CZK_HD static Fq13 mul(const Fq13& a, const Fq13& b) {
        uint64_t t[26];
#pragma unroll
        for (int k = 0; k < 26; k++) t[k] = 0;
#pragma unroll
        for (int i = 0; i < 13; i++) {
#pragma unroll
            for (int j = 0; j < 13; j++) t[i + j] += (uint64_t)a.d[i] * b.d[j];
            uint32_t m = ((uint32_t)t[i] * Fq13Params::INV) & MASK;
#pragma unroll
            for (int j = 0; j < 13; j++) t[i + j] += (uint64_t)m * Fq13Params::modc(j);
            t[i + 1] += t[i] >> 29;
        }
        Fq13 r;
#pragma unroll
        for (int k = 13; k < 25; k++) {
            t[k + 1] += t[k] >> 29;
            r.d[k - 13] = (uint32_t)t[k] & MASK;
        }
        r.d[12] = (uint32_t)t[25];  // value < 2p: top digit < 2^30
        cond_sub_p(r.d);
        return r;
    }
    CZK_HD static Fq13 sqr(const Fq13& a) { return mul(a, a); }
    CZK_HD_NOINLINE static Fq13 mul_ni(const Fq13& a, const Fq13& b) { return mul(a, b); }
    CZK_HD static Fq13 sqr_ni(const Fq13& a) { return mul_ni(a, a); }

    // 12 x 32-bit limbs <-> 13 x 29-bit digits of the same integer (< 2^377)
    CZK_HD static Fq13 reslice_from_limbs(const uint32_t* l) {
        Fq13 r;
#pragma unroll
        for (int k = 0; k < 13; k++) {
            int bit = 29 * k, w = bit >> 5, sh = bit & 31;
            uint64_t v = (uint64_t)l[w] | ((w + 1 < 12) ? ((uint64_t)l[w + 1] << 32) : 0ull);
            r.d[k] = (uint32_t)(v >> sh) & MASK;
        }
        return r;
    }
    CZK_HD void reslice_to_limbs(uint32_t* l) const {
#pragma unroll
        for (int w = 0; w < 12; w++) l[w] = 0;
#pragma unroll
        for (int k = 0; k < 13; k++) {
            int bit = 29 * k, w = bit >> 5, sh = bit & 31;
            uint64_t v = (uint64_t)d[k] << sh;
            l[w] |= (uint32_t)v;
            if (w + 1 < 12) l[w + 1] |= (uint32_t)(v >> 32);
        }
    }
    // reference form (x R mod p, R = 2^384, 12 limbs)  ->  digit form (x R13 mod p)
    CZK_HD static Fq13 from_std(const Fq& x) {
        Fq13 k;
#pragma unroll
        for (int i = 0; i < 13; i++) k.d[i] = Fq13Params::k_in(i);
        return mul(reslice_from_limbs(x.l), k);
    }
    CZK_HD Fq to_std() const {
        Fq13 k;
#pragma unroll
        for (int i = 0; i < 13; i++) k.d[i] = Fq13Params::k_out(i);
        Fq13 y = mul(*this, k);
        Fq r;
        y.reslice_to_limbs(r.l);
        return r;
    }
};

}  // namespace czk
