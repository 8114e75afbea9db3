// Modular inversion for the one-lane-per-block inversion of the batched-affine MSM rounds (msm_batched.cu, phase 2).
//
// The reference inverts with the binary extended Euclidean algorithm (algebra/ff/src/fields/macros.rs:368-422): up to
// 2 * 377 one-bit steps, every one of them a shift and a conditional add / subtract over all twelve limbs of four
// numbers - about 75 000 dependent instructions on a single GPU lane, which is what the other 127 lanes of the block wait
// for.  This is the same binary GCD with the one-bit steps batched 30 at a time (T. Pornin, "Optimized Binary GCD for
// Modular Inversion", ePrint 2020/972, algorithm 2): the 30 steps are run on 64-bit approximations of a and b (their
// low 30 bits and their top 32 bits, aligned to the longer of the two) while the update factors f0, g0, f1, g1
// (|f| + |g| <= 2^30: 32-bit registers) are collected, and only then applied to the full numbers
//     (a, b) <- ((a f0 + b g0) / 2^30, (a f1 + b g1) / 2^30)          exact divisions, signs fixed up afterwards
//     (u, v) <- ((u f0 + v g0) / 2^30, (u f1 + v g1) / 2^30)  mod p   one Montgomery-style reduction by 2^30 each
// 2 * 377 - 1 = 753 steps are enough for every input, i.e. 26 rounds of 30; one more round is run for margin (a round on
// a = 0, b = 1 changes nothing).  The invariants a = u * y / K and b = v * y / K (mod p) hold throughout, so starting
// from u = K = R^2 the result v = K / y is the inverse of a Montgomery-form input in Montgomery form - bit-identical to
// the reference's result, because a field element has one representation.  Plain 64-bit C++ (no carry intrinsics): it
// runs on one lane, what counts is the number of dependent steps, and the same source is checked on the host by
// tests/test_device_math_emulation.py.
#pragma once
#include "fp.cuh"

namespace czk {

CZK_HD int clz32(uint32_t x) {  // x != 0
#ifdef __CUDA_ARCH__
    return __clz((int)x);
#else
    return __builtin_clz(x);
#endif
}

template <class P>
struct BinGcd {
    static constexpr int N = P::N;
    static constexpr int STEPS = 30;                                    // binary-GCD steps per round
    static constexpr int ROUNDS = (2 * 32 * N - 1 + STEPS - 1) / STEPS + 1;  // ceil((2 bits - 1) / STEPS) + 1, bits <= 32 N
    static constexpr uint32_t LOW_MASK = (1u << STEPS) - 1;

    // r (N + 1 limbs) = x * k
    CZK_HD static void mul_word(uint32_t* r, const uint32_t* x, uint32_t k) {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            c += (uint64_t)x[i] * k;
            r[i] = (uint32_t)c;
            c >>= 32;
        }
        r[N] = (uint32_t)c;
    }
    // r (N + 1 limbs) += x * k; no carry out for the operand sizes used here
    CZK_HD static void mad_word(uint32_t* r, const uint32_t* x, uint32_t k) {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            c += (uint64_t)x[i] * k + r[i];
            r[i] = (uint32_t)c;
            c >>= 32;
        }
        r[N] += (uint32_t)c;
    }
    // out (N limbs) = r (N + 1 limbs) >> STEPS
    CZK_HD static void shr_steps(uint32_t* out, const uint32_t* r) {
#pragma unroll
        for (int i = 0; i < N; i++) out[i] = (r[i] >> STEPS) | (r[i + 1] << (32 - STEPS));
    }
    // |x fa +- y ga| / 2^STEPS into out; returns true when the signed value x (+-fa) + y (+-ga) is negative (or zero with
    // mixed signs, where the sign does not matter).  fneg / gneg: the signs of the two factors.
    CZK_HD static bool lincomb_exact(uint32_t* out, const uint32_t* x, uint32_t fa, bool fneg, const uint32_t* y, uint32_t ga,
                                     bool gneg) {
        uint32_t t[N + 1], s[N + 1];
        mul_word(t, x, fa);
        bool neg;
        if (fneg == gneg) {
            mad_word(t, y, ga);
            neg = fneg;
        } else {
            mul_word(s, y, ga);
            uint64_t br = 0;  // t -= s
#pragma unroll
            for (int i = 0; i <= N; i++) {
                uint64_t d = (uint64_t)t[i] - s[i] - br;
                t[i] = (uint32_t)d;
                br = (d >> 32) & 1u;
            }
            if (br) {  // t = -t
                uint64_t c = 1;
#pragma unroll
                for (int i = 0; i <= N; i++) {
                    c += (uint32_t)~t[i];
                    t[i] = (uint32_t)c;
                    c >>= 32;
                }
            }
            neg = fneg ? !br : (br != 0);
        }
        shr_steps(out, t);
        return neg;
    }
    // out = (x (+-fa) + y (+-ga)) / 2^STEPS mod p, for x, y in [0, p)
    CZK_HD static void lincomb_mod(uint32_t* out, const uint32_t* x, uint32_t fa, bool fneg, const uint32_t* y, uint32_t ga, bool gneg) {
        uint32_t xe[N], ye[N], m[N];
        {
            uint64_t bx = 0, by = 0;  // p - x and p - y (p itself when the operand is 0: still a representative of 0)
#pragma unroll
            for (int i = 0; i < N; i++) {
                m[i] = P::mod(i);
                uint64_t dx = (uint64_t)m[i] - x[i] - bx, dy = (uint64_t)m[i] - y[i] - by;
                xe[i] = fneg ? (uint32_t)dx : x[i];
                ye[i] = gneg ? (uint32_t)dy : y[i];
                bx = (dx >> 32) & 1u;
                by = (dy >> 32) & 1u;
            }
        }
        uint32_t t[N + 1];
        mul_word(t, xe, fa);
        mad_word(t, ye, ga);                       // <= p * 2^STEPS
        const uint32_t q = (0u - t[0]) & LOW_MASK;  // p = 1 mod 2^32: -p^-1 = -1 mod 2^STEPS
        mad_word(t, m, q);                         // = 0 mod 2^STEPS, < p * 2^(STEPS + 1)
        shr_steps(out, t);                         // < 2 p
        uint32_t d[N];
        uint64_t br = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            uint64_t e = (uint64_t)out[i] - m[i] - br;
            d[i] = (uint32_t)e;
            br = (e >> 32) & 1u;
        }
#pragma unroll
        for (int i = 0; i < N; i++) out[i] = br ? out[i] : d[i];
    }

    // y^-1 for y != 0, Montgomery form in and out
    CZK_HD static Fp<P> inverse(const Fp<P>& y) {
        static_assert(P::INV32 == 0xffffffffu, "the reduction by 2^STEPS above uses p = 1 mod 2^32");
        uint32_t a[N], b[N], u[N], v[N];
#pragma unroll
        for (int i = 0; i < N; i++) {
            a[i] = y.l[i];
            b[i] = P::mod(i);
            u[i] = P::r2(i);
            v[i] = 0;
        }
#pragma unroll 1
        for (int round = 0; round < ROUNDS; round++) {
            // 64-bit approximations: low STEPS bits | top 32 bits of the longer of a, b (exact when both fit 64 bits)
            uint32_t ah = a[1], al = a[0], bh = b[1], bl = b[0];
            bool found = false;
#pragma unroll
            for (int i = N - 1; i >= 2; i--) {
                const bool hit = !found && (a[i] | b[i]) != 0;
                ah = hit ? a[i] : ah;
                al = hit ? a[i - 1] : al;
                bh = hit ? b[i] : bh;
                bl = hit ? b[i - 1] : bl;
                found = found || hit;
            }
            uint64_t xa = ((uint64_t)ah << 32) | al, xb = ((uint64_t)bh << 32) | bl;
            if (found) {
                const int lz = clz32(ah | bh);
                // the pair (top word, next word) holds 64 - lz significant bits of the longer number: keep its top 32
                xa = ((xa >> (32 - lz)) << STEPS) | (a[0] & LOW_MASK);
                xb = ((xb >> (32 - lz)) << STEPS) | (b[0] & LOW_MASK);
            }
            // STEPS binary-GCD steps on the approximations, collecting the factors
            int32_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
#pragma unroll 1
            for (int j = 0; j < STEPS; j++) {
                const bool odd = xa & 1u;
                const bool swap = odd && xa < xb;
                const uint64_t ta = swap ? xb : xa, tb = swap ? xa : xb;
                const int32_t tf0 = swap ? f1 : f0, tf1 = swap ? f0 : f1, tg0 = swap ? g1 : g0, tg1 = swap ? g0 : g1;
                xa = (odd ? ta - tb : ta) >> 1;
                xb = tb;
                f0 = odd ? tf0 - tf1 : tf0;
                g0 = odd ? tg0 - tg1 : tg0;
                f1 = tf1 * 2;
                g1 = tg1 * 2;
            }
            // apply them: |f| + |g| <= 2^STEPS
            bool fn0 = f0 < 0, gn0 = g0 < 0, fn1 = f1 < 0, gn1 = g1 < 0;
            const uint32_t fa0 = (uint32_t)(fn0 ? -f0 : f0), ga0 = (uint32_t)(gn0 ? -g0 : g0);
            const uint32_t fa1 = (uint32_t)(fn1 ? -f1 : f1), ga1 = (uint32_t)(gn1 ? -g1 : g1);
            uint32_t na[N], nb[N];
            const bool nega = lincomb_exact(na, a, fa0, fn0, b, ga0, gn0);
            const bool negb = lincomb_exact(nb, a, fa1, fn1, b, ga1, gn1);
            fn0 ^= nega;
            gn0 ^= nega;
            fn1 ^= negb;
            gn1 ^= negb;
            uint32_t nu[N], nv[N];
            lincomb_mod(nu, u, fa0, fn0, v, ga0, gn0);
            lincomb_mod(nv, u, fa1, fn1, v, ga1, gn1);
#pragma unroll
            for (int i = 0; i < N; i++) {
                a[i] = na[i];
                b[i] = nb[i];
                u[i] = nu[i];
                v[i] = nv[i];
            }
        }
        Fp<P> r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = v[i];
        return r;
    }
};

}  // namespace czk
