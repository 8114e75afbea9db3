// Montgomery prime-field arithmetic on 32-bit limbs held in registers.
//
// Values are bit-identical to the reference's Fp256/Fp384 Montgomery
// representation (algebra/ff/src/fields/macros.rs:89-703, R = 2^256 / 2^384):
// N 32-bit limbs little-endian == N/2 of the reference's u64 limbs, always fully
// reduced to [0, p) on return, so device buffers can be memcpy'd to and from the
// reference's in-memory field elements.
//
// mul() is an operand-scanning Montgomery product (the same CIOS recurrence as
// algebra/ff/src/fields/arithmetic.rs:36-52, on 32-bit words) scheduled so that
// every 32x32->64 product lands on an even-aligned register pair: the partial
// sums are kept in two interleaved accumulators, one for the products whose
// column index is even, one for the odd columns, each a carry chain of
// IMAD.WIDE.U32.X.  Dividing by 2^32 after each row is free - the two
// accumulators just swap roles.  2*N^2 wide multiply-adds per product (288 for
// Fq, 128 for Fr).
#pragma once
#include "carry.cuh"
#include "bls12_377_params.cuh"

namespace czk {

// acc (N words, N/2 aligned pairs) += a[0], a[2], a[4], ... times bi, carries chained pair to pair.
// `a` may be offset by one to address the odd limbs.  The carry out of the top pair is left in CF.
// bi2 == bi in value.  Passing the same register lets ptxas fuse each (lo, hi) pair into IMAD.WIDE.U32.X
// (measured 19.9 lanes/clk/SM on B200); passing an opaque copy keeps them apart as IMAD + IMAD.HI.U32 on the
// multiply pipe with the carries as IADD3.X on the ALU pipe, which overlaps the two pipes.
template <int N>
CZK_HD void chain_mad(uint32_t* acc, const uint32_t* a, uint32_t bi, uint32_t bi2) {
    acc[0] = mad_lo_cc(a[0], bi, acc[0]);
    acc[1] = madc_hi_cc(a[0], bi2, acc[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        acc[j] = madc_lo_cc(a[j], bi, acc[j]);
        acc[j + 1] = madc_hi_cc(a[j], bi2, acc[j + 1]);
    }
}
template <int N>
CZK_HD void chain_mad(uint32_t* acc, const uint32_t* a, uint32_t bi) {
    chain_mad<N>(acc, a, bi, bi);
}

#ifdef __CUDACC__
// 0xffffffff from constant memory: a value ptxas cannot see.  Both BLS12-377 moduli are 1 mod 2^32, so the Montgomery
// factor m = -t0 mod 2^32; written as a visible negation ptxas refuses to fuse the (mad.lo.cc, madc.hi.cc) pairs
// that multiply by it into IMAD.WIDE.U32.X and emits IMAD.X + IMAD.HI.U32.X instead (twice the multiply-pipe
// slots for the whole reduction half of the product).  (t0 ^ ones) + 1 is the same value, opaque.
static __device__ __constant__ uint32_t FP_ALL_ONES_C = 0xffffffffu;
#endif
template <class P>
CZK_HD uint32_t mont_factor(uint32_t t0) {
    static_assert(P::INV32 == 0xffffffffu, "both moduli are 1 mod 2^32");
#ifdef __CUDA_ARCH__
    return (t0 ^ FP_ALL_ONES_C) + 1u;
#else
    return t0 * P::INV32;
#endif
}

// One row of the product: add a*bi and the Montgomery multiple of p that clears column 0.
//   E holds columns 0..N-1 (pairs at even columns), O holds columns 1..N (pairs at odd columns).
// On entry (not first) a division by 2^32 from the previous row is still pending, which is why
// the caller swaps E and O between rows: old O *is* the new E, and old E shifted down two words
// is the new O - that shift is folded into the multiply-adds that refill O.
// The two halves of a row.  mont_row_acc adds a * bi (and performs the pending division by 2^32 of the previous row, see
// above); mont_row_red adds the Montgomery multiple of p that clears column 0.
template <class P>
CZK_HD void mont_row_acc(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi, bool first, uint32_t opaque_zero = 0) {
    constexpr int N = P::N;
    const uint32_t bi2 = bi + opaque_zero;
    if (first) {
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            O[j] = mul_lo(a[j + 1], bi);
            O[j + 1] = mul_hi(a[j + 1], bi);
        }
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            E[j] = mul_lo(a[j], bi);
            E[j + 1] = mul_hi(a[j], bi);
        }
    } else {
        E[0] = add_cc(E[0], O[1]);
#pragma unroll
        for (int j = 0; j < N - 2; j += 2) {
            O[j] = madc_lo_cc(a[j + 1], bi, O[j + 2]);
            O[j + 1] = madc_hi_cc(a[j + 1], bi2, O[j + 3]);
        }
        O[N - 2] = madc_lo_cc(a[N - 1], bi, 0);
        O[N - 1] = madc_hi(a[N - 1], bi2, 0);
        chain_mad<N>(E, a, bi, bi2);
        O[N - 1] = addc(O[N - 1], 0);
    }
}
// a second product accumulated into the same row: E, O += a2 * b2i (no shift: mont_row_acc already did it).  The caller
// guarantees that the running total stays below 2^(32 (N + 1)) (see Fp::mul_sum2).
template <class P>
CZK_HD void mont_row_acc2(uint32_t* E, uint32_t* O, const uint32_t* a2, uint32_t b2i) {
    constexpr int N = P::N;
    chain_mad<N>(O, a2 + 1, b2i);  // columns 1..N; no carry out of column N by the caller's bound
    chain_mad<N>(E, a2, b2i);      // columns 0..N-1; the carry out lands in column N
    O[N - 1] = addc(O[N - 1], 0);
}
template <class P>
CZK_HD void mont_row_red(uint32_t* E, uint32_t* O, const uint32_t* m) {
    constexpr int N = P::N;
    uint32_t mi = mont_factor<P>(E[0]);
    chain_mad<N>(O, m + 1, mi);
    // E += mi * (p0, p2, p4, ...): p0 = 1, so the first product is mi itself (no multiply): E[0] + mi = 0 mod 2^32
    static_assert(P::mod(0) == 1u, "both moduli are 1 mod 2^32");
    E[0] = add_cc(E[0], mi);
    E[1] = addc_cc(E[1], 0);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        E[j] = madc_lo_cc(m[j], mi, E[j]);
        E[j + 1] = madc_hi_cc(m[j], mi, E[j + 1]);
    }
    O[N - 1] = addc(O[N - 1], 0);
}
template <class P>
CZK_HD void mont_row(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi, const uint32_t* m, bool first,
                     uint32_t opaque_zero = 0) {
    mont_row_acc<P>(E, O, a, bi, first, opaque_zero);
    mont_row_red<P>(E, O, m);
}

template <class P>
struct Fp {
    static constexpr int N = P::N;
    uint32_t l[N];

    CZK_HD static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    CZK_HD static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::one(i);
        return r;
    }
    CZK_HD static Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::r2(i);
        return r;
    }
    CZK_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= l[i];
        return o == 0;
    }
    CZK_HD bool operator==(const Fp& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= l[i] ^ b.l[i];
        return o == 0;
    }
    CZK_HD bool operator!=(const Fp& b) const { return !(*this == b); }

    // r = x - p if x >= p else x   (x < 2p)
    CZK_HD static void reduce_once(uint32_t* x) {
        uint32_t t[N];
        t[0] = sub_cc(x[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = subc_cc(x[i], P::mod(i));
        uint32_t borrow = subc(0, 0);  // 0xffffffff if x < p
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = borrow ? x[i] : t[i];
    }

    CZK_HD static Fp add(const Fp& a, const Fp& b) {
        Fp r;
        r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
        r.l[N - 1] = addc(a.l[N - 1], b.l[N - 1]);
        reduce_once(r.l);
        return r;
    }
    CZK_HD static Fp sub(const Fp& a, const Fp& b) {
        Fp r;
        r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) r.l[i] = subc_cc(a.l[i], b.l[i]);
        uint32_t borrow = subc(0, 0);  // all ones if a < b
        r.l[0] = add_cc(r.l[0], P::mod(0) & borrow);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(r.l[i], P::mod(i) & borrow);
        r.l[N - 1] = addc(r.l[N - 1], P::mod(N - 1) & borrow);
        return r;
    }
    CZK_HD static Fp dbl(const Fp& a) { return add(a, a); }
    CZK_HD static Fp neg(const Fp& a) {
        Fp r;
        r.l[0] = sub_cc(P::mod(0), a.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = subc_cc(P::mod(i), a.l[i]);
        r.l[N - 1] = subc(P::mod(N - 1), a.l[N - 1]);
        uint32_t nz = a.is_zero() ? 0u : 0xffffffffu;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] &= nz;
        return r;
    }
    CZK_HD static Fp mul(const Fp& a, const Fp& b) {
        uint32_t m[N];
#pragma unroll
        for (int i = 0; i < N; i++) m[i] = P::modc(i);
        return mul_m(a, b, m);
    }
    // the product with the modulus limbs supplied by the caller (kernels that keep p in registers, loaded from
    // memory the compiler cannot see through, get every reduction row fused into IMAD.WIDE.U32.X)
    CZK_HD static Fp mul_m(const Fp& a, const Fp& b, const uint32_t* m) {
        uint32_t even[N], odd[N];
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            mont_row<P>(even, odd, a.l, b.l[i], m, i == 0);
            mont_row<P>(odd, even, a.l, b.l[i + 1], m, false);
        }
        Fp r;
        r.l[0] = add_cc(even[0], odd[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(even[i], odd[i + 1]);
        r.l[N - 1] = addc(even[N - 1], 0);
        reduce_once(r.l);
        return r;
    }
    // (a * b + c * d) / R mod p with ONE Montgomery reduction for the two products (lazy reduction: the rows of both
    // products are accumulated before each reduction row): 3 N^2 - N multiply-adds instead of 4 N^2 - 2 N.
    // a, b, d < p; c may be an UNREDUCED N-limb value below 6 p (Fq2 passes -5 x as 5 (p - x)).  Bounds: every row keeps
    // the running total below (previous / 2^32) + a + c + p < 8 p < 2^(32 N) (p < 2^(32 N - 7) for both fields), and
    // the result is below (a b + c d) / R + p < p (7 p / R + 1) < 2 p, so one conditional subtraction reduces it.
    CZK_HD static Fp mul_sum2(const Fp& a, const Fp& b, const uint32_t* c, const Fp& d) {
        uint32_t m[N], even[N], odd[N];
#pragma unroll
        for (int i = 0; i < N; i++) m[i] = P::modc(i);
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            mont_row_acc<P>(even, odd, a.l, b.l[i], i == 0);
            mont_row_acc2<P>(even, odd, c, d.l[i]);
            mont_row_red<P>(even, odd, m);
            mont_row_acc<P>(odd, even, a.l, b.l[i + 1], false);
            mont_row_acc2<P>(odd, even, c, d.l[i + 1]);
            mont_row_red<P>(odd, even, m);
        }
        Fp r;
        r.l[0] = add_cc(even[0], odd[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(even[i], odd[i + 1]);
        r.l[N - 1] = addc(even[N - 1], 0);
        reduce_once(r.l);
        return r;
    }
    // mul_sum2 with the multiplier limbs produced on demand: fbd(i, b_i, d_i) (the two-lane Fq2 product fetches
    // them from the partner lane row by row instead of holding two more N-limb operands in registers)
    template <class FBD>
    CZK_HD static Fp mul_sum2_f(const Fp& a, const uint32_t* c, FBD fbd) {
        uint32_t m[N], even[N], odd[N];
#pragma unroll
        for (int i = 0; i < N; i++) m[i] = P::modc(i);
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            uint32_t bi, di;
            fbd(i, bi, di);
            mont_row_acc<P>(even, odd, a.l, bi, i == 0);
            mont_row_acc2<P>(even, odd, c, di);
            mont_row_red<P>(even, odd, m);
            fbd(i + 1, bi, di);
            mont_row_acc<P>(odd, even, a.l, bi, false);
            mont_row_acc2<P>(odd, even, c, di);
            mont_row_red<P>(odd, even, m);
        }
        Fp r;
        r.l[0] = add_cc(even[0], odd[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(even[i], odd[i + 1]);
        r.l[N - 1] = addc(even[N - 1], 0);
        reduce_once(r.l);
        return r;
    }
    // 5 (p - a) as an unreduced N-limb integer (< 5 p + 1 <= 2^(32 N - 4)): -5 a mod p for mul_sum2's `c` operand
    CZK_HD static void neg_times5_unreduced(const Fp& a, uint32_t* out) {
        uint32_t t[N];
        t[0] = sub_cc(P::mod(0), a.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) t[i] = subc_cc(P::mod(i), a.l[i]);
        t[N - 1] = subc(P::mod(N - 1), a.l[N - 1]);
        // 4 t + t
        uint32_t q[N];
        q[0] = t[0] << 2;
#pragma unroll
        for (int i = 1; i < N; i++) q[i] = (t[i] << 2) | (t[i - 1] >> 30);
        out[0] = add_cc(q[0], t[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) out[i] = addc_cc(q[i], t[i]);
        out[N - 1] = addc(q[N - 1], t[N - 1]);
    }
    // Two independent products with their rows interleaved in program order (the carry primitives are `asm volatile`,
    // so two plain mul() calls stay one after the other and ptxas never overlaps them): four carry chains in flight
    // instead of two.  An experiment for kernels that run few warps per scheduler (microbench kind 15); 2 x 26
    // accumulator registers.
    CZK_HD static void mul2(const Fp& a, const Fp& b, const Fp& c, const Fp& d, Fp& ab, Fp& cd) {
        uint32_t m[N], e1[N], o1[N], e2[N], o2[N];
#pragma unroll
        for (int i = 0; i < N; i++) m[i] = P::modc(i);
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            mont_row<P>(e1, o1, a.l, b.l[i], m, i == 0);
            mont_row<P>(e2, o2, c.l, d.l[i], m, i == 0);
            mont_row<P>(o1, e1, a.l, b.l[i + 1], m, false);
            mont_row<P>(o2, e2, c.l, d.l[i + 1], m, false);
        }
        ab.l[0] = add_cc(e1[0], o1[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) ab.l[i] = addc_cc(e1[i], o1[i + 1]);
        ab.l[N - 1] = addc(e1[N - 1], 0);
        cd.l[0] = add_cc(e2[0], o2[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) cd.l[i] = addc_cc(e2[i], o2[i + 1]);
        cd.l[N - 1] = addc(e2[N - 1], 0);
        reduce_once(ab.l);
        reduce_once(cd.l);
    }
    CZK_HD static Fp sqr(const Fp& a) { return mul(a, a); }
    // out-of-line copy: one body per kernel instead of one per call site.  Used wherever code size
    // (compile time, instruction cache) matters more than the call overhead: Fq2 and the cold kernels.
    CZK_HD_NOINLINE static Fp mul_ni(const Fp& a, const Fp& b) { return mul(a, b); }
    CZK_HD static Fp sqr_ni(const Fp& a) { return mul_ni(a, a); }

    // Montgomery form -> canonical integer (arithmetic.rs:59-82): multiply by 1.
    CZK_HD static Fp from_mont(const Fp& a) {
        Fp o = zero();
        o.l[0] = 1;
        return mul(a, o);
    }
    // canonical integer -> Montgomery form (macros.rs:444-454): multiply by R^2.
    CZK_HD static Fp to_mont(const Fp& a) { return mul(a, r2()); }

    // a^e for a small public exponent (square and multiply, MSB first)
    CZK_HD static Fp pow_u64(const Fp& a, uint64_t e) {
        Fp res = one();
        bool started = false;
        for (int i = 63; i >= 0; i--) {
            if (started) res = sqr(res);
            if ((e >> i) & 1) {
                started = true;
                res = mul(res, a);
            }
        }
        return res;
    }
    // a^(p-2): inversion for the places the device needs one (batch normalisation of synthetic
    // bases).  The hot paths never invert on the device.
    CZK_HD static Fp inv_fermat(const Fp& a) {
        Fp res = one();
        // exponent p-2, limbs from P::mod with the borrow handled explicitly (p is odd, p0 >= 3 or
        // low limb 1 -> borrow); scan from the top bit
        uint32_t e[N];
        {
            e[0] = sub_cc(P::mod(0), 2);
#pragma unroll
            for (int i = 1; i < N; i++) e[i] = subc_cc(P::mod(i), 0);
        }
        bool started = false;
        for (int i = N * 32 - 1; i >= 0; i--) {
            if (started) res = sqr(res);
            if ((e[i >> 5] >> (i & 31)) & 1) {
                started = true;
                res = mul(res, a);
            }
        }
        return res;
    }
};

using Fr = Fp<FrParams>;
using Fq = Fp<FqParams>;

// Fq whose products are calls to ONE out-of-line body instead of ten inlined copies per point addition
// (smaller instruction footprint; same limbs, same results).
struct FqCall : Fq {
    CZK_HD FqCall() {}
    CZK_HD FqCall(const Fq& f) : Fq(f) {}
    CZK_HD static FqCall zero() { return FqCall(Fq::zero()); }
    CZK_HD static FqCall one() { return FqCall(Fq::one()); }
    CZK_HD static FqCall add(const FqCall& a, const FqCall& b) { return FqCall(Fq::add(a, b)); }
    CZK_HD static FqCall sub(const FqCall& a, const FqCall& b) { return FqCall(Fq::sub(a, b)); }
    CZK_HD static FqCall dbl(const FqCall& a) { return FqCall(Fq::dbl(a)); }
    CZK_HD static FqCall neg(const FqCall& a) { return FqCall(Fq::neg(a)); }
    CZK_HD static FqCall mul(const FqCall& a, const FqCall& b) { return FqCall(Fq::mul_ni(a, b)); }
    CZK_HD static FqCall sqr(const FqCall& a) { return FqCall(Fq::mul_ni(a, a)); }
};

// Fq2 = Fq[u]/(u^2 + 5)  (curves/bls12_377/src/fields/fq2.rs:13; NONRESIDUE = -5)
struct Fq2 {
    Fq c0, c1;
    CZK_HD static Fq2 zero() { return Fq2{Fq::zero(), Fq::zero()}; }
    CZK_HD static Fq2 one() { return Fq2{Fq::one(), Fq::zero()}; }
    CZK_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    CZK_HD bool operator==(const Fq2& b) const { return c0 == b.c0 && c1 == b.c1; }
    CZK_HD bool operator!=(const Fq2& b) const { return !(*this == b); }
    CZK_HD static Fq2 add(const Fq2& a, const Fq2& b) { return Fq2{Fq::add(a.c0, b.c0), Fq::add(a.c1, b.c1)}; }
    CZK_HD static Fq2 sub(const Fq2& a, const Fq2& b) { return Fq2{Fq::sub(a.c0, b.c0), Fq::sub(a.c1, b.c1)}; }
    CZK_HD static Fq2 dbl(const Fq2& a) { return Fq2{Fq::dbl(a.c0), Fq::dbl(a.c1)}; }
    CZK_HD static Fq2 neg(const Fq2& a) { return Fq2{Fq::neg(a.c0), Fq::neg(a.c1)}; }
    // -5 * x
    CZK_HD static Fq mul_by_nonresidue(const Fq& x) {
        Fq x2 = Fq::dbl(x);
        Fq x4 = Fq::dbl(x2);
        return Fq::neg(Fq::add(x4, x));
    }
    // Karatsuba, 3 base-field products (quadratic_extension.rs:569-583)
    CZK_HD static Fq2 mul_inl(const Fq2& a, const Fq2& b) {
        Fq v0 = Fq::mul(a.c0, b.c0);
        Fq v1 = Fq::mul(a.c1, b.c1);
        Fq t = Fq::mul(Fq::add(a.c0, a.c1), Fq::add(b.c0, b.c1));
        t = Fq::sub(Fq::sub(t, v0), v1);
        return Fq2{Fq::add(v0, mul_by_nonresidue(v1)), t};
    }
    // (c0^2 - 5 c1^2, 2 c0 c1) with 2 base-field products (quadratic_extension.rs:257-306)
    CZK_HD static Fq2 sqr_inl(const Fq2& a) {
        Fq v0 = Fq::sub(a.c0, a.c1);
        Fq v3 = Fq::sub(a.c0, mul_by_nonresidue(a.c1));
        Fq v2 = Fq::mul(a.c0, a.c1);
        v0 = Fq::mul(v0, v3);
        Fq c1 = Fq::dbl(v2);
        Fq c0 = Fq::add(Fq::add(v0, v2), mul_by_nonresidue(v2));
        return Fq2{c0, c1};
    }
    // out-of-line bodies: what the long point-addition formulas call (operands and result travel through local memory)
    CZK_HD_NOINLINE static Fq2 mul(const Fq2& a, const Fq2& b) { return mul_inl(a, b); }
    CZK_HD_NOINLINE static Fq2 sqr(const Fq2& a) { return sqr_inl(a); }
    CZK_HD_NOINLINE static Fq2 inv_fermat(const Fq2& a) {
        // 1/(c0 + c1 u) = (c0 - c1 u) / (c0^2 + 5 c1^2)
        Fq norm = Fq::sub(Fq::sqr(a.c0), mul_by_nonresidue(Fq::sqr(a.c1)));
        Fq ni = Fq::inv_fermat(norm);
        return Fq2{Fq::mul(a.c0, ni), Fq::neg(Fq::mul(a.c1, ni))};
    }
};

}  // namespace czk
