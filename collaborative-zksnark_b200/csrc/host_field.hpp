// Host-side (CPU) field and group arithmetic used by the library's glue code:
// domain constants, the O(windows) tail of an MSM (window combine + one affine
// normalisation), and the handful of group operations of the Groth16 prover that
// are not data parallel (r*s*delta, adding vk elements, share bookkeeping).
// 64-bit limbs, unsigned __int128 products.  This is product code (it ships in
// libczk_b200.so); it is independent of oracle/.
#pragma once
#include <cstdint>
#include <cstring>
#include "bls12_377_params.cuh"

namespace czk {
namespace host {

typedef unsigned __int128 u128;

template <class P>
struct HFp {
    static constexpr int N = P::N64;
    uint64_t l[N];

    static HFp zero() {
        HFp r;
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    static HFp one() {
        HFp r;
        for (int i = 0; i < N; i++) r.l[i] = P::ONE64[i];
        return r;
    }
    static HFp from_limbs(const uint64_t* p) {
        HFp r;
        std::memcpy(r.l, p, sizeof r.l);
        return r;
    }
    void to_limbs(uint64_t* p) const { std::memcpy(p, l, sizeof l); }
    bool is_zero() const {
        uint64_t o = 0;
        for (int i = 0; i < N; i++) o |= l[i];
        return o == 0;
    }
    bool operator==(const HFp& b) const {
        uint64_t o = 0;
        for (int i = 0; i < N; i++) o |= l[i] ^ b.l[i];
        return o == 0;
    }
    bool operator!=(const HFp& b) const { return !(*this == b); }
    static bool geq_mod(const uint64_t* a) {
        for (int i = N - 1; i >= 0; i--) {
            if (a[i] > P::MOD64[i]) return true;
            if (a[i] < P::MOD64[i]) return false;
        }
        return true;
    }
    static void sub_mod(uint64_t* a) {
        uint64_t borrow = 0;
        for (int i = 0; i < N; i++) {
            u128 d = (u128)a[i] - P::MOD64[i] - borrow;
            a[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
    }
    static HFp add(const HFp& a, const HFp& b) {
        HFp r;
        u128 c = 0;
        for (int i = 0; i < N; i++) {
            c += (u128)a.l[i] + b.l[i];
            r.l[i] = (uint64_t)c;
            c >>= 64;
        }
        if (geq_mod(r.l)) sub_mod(r.l);
        return r;
    }
    static HFp sub(const HFp& a, const HFp& b) {
        HFp r;
        uint64_t borrow = 0;
        for (int i = 0; i < N; i++) {
            u128 d = (u128)a.l[i] - b.l[i] - borrow;
            r.l[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
        if (borrow) {
            u128 c = 0;
            for (int i = 0; i < N; i++) {
                c += (u128)r.l[i] + P::MOD64[i];
                r.l[i] = (uint64_t)c;
                c >>= 64;
            }
        }
        return r;
    }
    static HFp dbl(const HFp& a) { return add(a, a); }
    static HFp neg(const HFp& a) { return a.is_zero() ? a : sub(zero(), a); }
    // word-serial Montgomery product (separate accumulate / reduce rows, N+2 word accumulator)
    static HFp mul(const HFp& a, const HFp& b) {
        uint64_t t[N + 2];
        for (int i = 0; i < N + 2; i++) t[i] = 0;
        for (int i = 0; i < N; i++) {
            u128 c = 0;
            for (int j = 0; j < N; j++) {
                c += (u128)a.l[j] * b.l[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[N];
            t[N] = (uint64_t)c;
            t[N + 1] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * P::INV64;
            c = (u128)m * P::MOD64[0] + t[0];
            c >>= 64;
            for (int j = 1; j < N; j++) {
                c += (u128)m * P::MOD64[j] + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[N];
            t[N - 1] = (uint64_t)c;
            t[N] = t[N + 1] + (uint64_t)(c >> 64);
        }
        HFp r;
        for (int i = 0; i < N; i++) r.l[i] = t[i];
        if (t[N] || geq_mod(r.l)) sub_mod(r.l);
        return r;
    }
    static HFp sqr(const HFp& a) { return mul(a, a); }
    static HFp from_u64(uint64_t x) {
        HFp r = zero();
        r.l[0] = x;
        HFp r2;
        for (int i = 0; i < N; i++) r2.l[i] = P::R2_64[i];
        return mul(r, r2);
    }
    HFp from_mont() const {
        HFp o = zero();
        o.l[0] = 1;
        return mul(*this, o);
    }
    HFp to_mont() const {
        HFp r2;
        for (int i = 0; i < N; i++) r2.l[i] = P::R2_64[i];
        return mul(*this, r2);
    }
    static HFp pow(const HFp& a, const uint64_t* e, int ne) {
        HFp res = one();
        bool started = false;
        for (int i = ne * 64 - 1; i >= 0; i--) {
            if (started) res = sqr(res);
            if ((e[i / 64] >> (i % 64)) & 1) {
                started = true;
                res = mul(res, a);
            }
        }
        return res;
    }
    static HFp pow_u64(const HFp& a, uint64_t e) { return pow(a, &e, 1); }
    // a^(p-2)
    static HFp inv(const HFp& a) {
        uint64_t e[N];
        for (int i = 0; i < N; i++) e[i] = P::MOD64[i];
        // p - 2 (p is odd and its low limb is >= 3 for both fields)
        e[0] -= 2;
        return pow(a, e, N);
    }
};

using HFr = HFp<FrParams>;
using HFq = HFp<FqParams>;

struct HFq2 {
    HFq c0, c1;
    static HFq2 zero() { return HFq2{HFq::zero(), HFq::zero()}; }
    static HFq2 one() { return HFq2{HFq::one(), HFq::zero()}; }
    static HFq2 from_limbs(const uint64_t* p) { return HFq2{HFq::from_limbs(p), HFq::from_limbs(p + 6)}; }
    void to_limbs(uint64_t* p) const {
        c0.to_limbs(p);
        c1.to_limbs(p + 6);
    }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const HFq2& b) const { return c0 == b.c0 && c1 == b.c1; }
    bool operator!=(const HFq2& b) const { return !(*this == b); }
    static HFq2 add(const HFq2& a, const HFq2& b) { return HFq2{HFq::add(a.c0, b.c0), HFq::add(a.c1, b.c1)}; }
    static HFq2 sub(const HFq2& a, const HFq2& b) { return HFq2{HFq::sub(a.c0, b.c0), HFq::sub(a.c1, b.c1)}; }
    static HFq2 dbl(const HFq2& a) { return add(a, a); }
    static HFq2 neg(const HFq2& a) { return HFq2{HFq::neg(a.c0), HFq::neg(a.c1)}; }
    static HFq nr(const HFq& x) {  // -5 x
        HFq x4 = HFq::dbl(HFq::dbl(x));
        return HFq::neg(HFq::add(x4, x));
    }
    static HFq2 mul(const HFq2& a, const HFq2& b) {
        HFq v0 = HFq::mul(a.c0, b.c0), v1 = HFq::mul(a.c1, b.c1);
        HFq t = HFq::mul(HFq::add(a.c0, a.c1), HFq::add(b.c0, b.c1));
        return HFq2{HFq::add(v0, nr(v1)), HFq::sub(HFq::sub(t, v0), v1)};
    }
    static HFq2 sqr(const HFq2& a) { return mul(a, a); }
    static HFq2 inv(const HFq2& a) {
        HFq norm = HFq::sub(HFq::sqr(a.c0), nr(HFq::sqr(a.c1)));
        HFq ni = HFq::inv(norm);
        return HFq2{HFq::mul(a.c0, ni), HFq::neg(HFq::mul(a.c1, ni))};
    }
};

// XYZZ group law on the host (same formulas as ec.cuh)
template <class F>
struct HPoint {
    F x, y, zz, zzz;
    static HPoint infinity() { return HPoint{F::zero(), F::zero(), F::zero(), F::zero()}; }
    static HPoint from_affine(const F& ax, const F& ay) { return HPoint{ax, ay, F::one(), F::one()}; }
    bool is_inf() const { return zz.is_zero(); }
    static HPoint dbl(const HPoint& p) {
        if (p.is_inf()) return p;
        F u = F::dbl(p.y), v = F::sqr(u), w = F::mul(u, v), s = F::mul(p.x, v);
        F xx = F::sqr(p.x), m = F::add(F::dbl(xx), xx);
        HPoint r;
        r.x = F::sub(F::sub(F::sqr(m), s), s);
        r.y = F::sub(F::mul(m, F::sub(s, r.x)), F::mul(w, p.y));
        r.zz = F::mul(v, p.zz);
        r.zzz = F::mul(w, p.zzz);
        return r;
    }
    void add(const HPoint& o) {
        if (o.is_inf()) return;
        if (is_inf()) {
            *this = o;
            return;
        }
        F u1 = F::mul(x, o.zz), u2 = F::mul(o.x, zz), s1 = F::mul(y, o.zzz), s2 = F::mul(o.y, zzz);
        F p = F::sub(u2, u1), r = F::sub(s2, s1);
        if (p.is_zero() && r.is_zero()) {
            *this = dbl(*this);
            return;
        }
        F pp = F::sqr(p), ppp = F::mul(p, pp), q = F::mul(u1, pp);
        F x3 = F::sub(F::sub(F::sub(F::sqr(r), ppp), q), q);
        F y3 = F::sub(F::mul(r, F::sub(q, x3)), F::mul(s1, ppp));
        x = x3;
        y = y3;
        zz = F::mul(F::mul(zz, o.zz), pp);
        zzz = F::mul(F::mul(zzz, o.zzz), ppp);
    }
    void negate() { y = F::neg(y); }
    // scalar given as canonical little-endian u64 limbs
    // k * p, k = nk little-endian 64-bit limbs: fixed 4-bit windows over a table of 1p .. 15p
    static HPoint mul(const HPoint& p, const uint64_t* k, int nk) {
        int top = nk * 64 - 1;
        while (top >= 0 && !((k[top / 64] >> (top % 64)) & 1)) top--;
        if (top < 0) return infinity();
        HPoint tab[16];
        tab[0] = infinity();
        tab[1] = p;
        for (int i = 2; i < 16; i++) {
            tab[i] = (i & 1) ? tab[i - 1] : dbl(tab[i / 2]);
            if (i & 1) tab[i].add(p);
        }
        HPoint acc = infinity();
        for (int w = top / 4; w >= 0; w--) {
            if (!acc.is_inf())
                for (int j = 0; j < 4; j++) acc = dbl(acc);
            const unsigned d = (unsigned)(k[w / 16] >> (4 * (w % 16))) & 15u;
            if (d) acc.add(tab[d]);
        }
        return acc;
    }
    // -> affine; returns false for infinity
    bool to_affine(F& ax, F& ay) const {
        if (is_inf()) return false;
        // 1/zz and 1/zzz from one inversion: inv = 1/(zz*zzz)
        F inv = F::inv(F::mul(zz, zzz));
        ax = F::mul(x, F::mul(inv, zzz));
        ay = F::mul(y, F::mul(inv, zz));
        return true;
    }
};

using HG1 = HPoint<HFq>;
using HG2 = HPoint<HFq2>;

}  // namespace host
}  // namespace czk
