// Short-Weierstrass (a = 0) group law in extended Jacobian ("XYZZ") coordinates,
// generic over the base field (Fq for G1, Fq2 for G2).
//
// The reference accumulates buckets in Jacobian coordinates
// (algebra/ec/src/models/short_weierstrass_jacobian.rs:570-638, madd-2007-bl, 7M+4S
// and a field doubling chain).  A projective representative is not canonical, so
// the device is free to use XYZZ (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; infinity <=> ZZ = 0):
// mixed addition is 8M+2S with no doublings of field elements, and the reference's
// three special cases keep their meaning:
//   P + inf, inf + P                       (:571-580)
//   P + P  -> doubling                     (:594-596)
//   P + -P -> infinity (falls out of the formula: ZZ3 = ZZ1 * 0)   (:598)
// Only the affine value of the final sum is compared with the reference.
#pragma once
#include "fp.cuh"

namespace czk {

template <class F>
struct Affine {
    F x, y;  // (0, 0) is never on y^2 = x^3 + b (b != 0): used as the in-kernel infinity marker
};

template <class F>
struct XYZZ {
    F x, y, zz, zzz;

    CZK_HD static XYZZ infinity() { return XYZZ{F::zero(), F::zero(), F::zero(), F::zero()}; }
    CZK_HD bool is_inf() const { return zz.is_zero(); }

    CZK_HD static XYZZ from_affine(const F& ax, const F& ay) { return XYZZ{ax, ay, F::one(), F::one()}; }

    // 2 * (ax, ay) for an affine point (mdbl-2008-s-1 with a = 0).  ay = 0 cannot happen in the
    // prime-order subgroup; if it does V = 0 and the result is the infinity encoding, which is right.
    CZK_HD_NOINLINE static XYZZ dbl_affine(const F& ax, const F& ay) {
        F u = F::dbl(ay);
        F v = F::sqr(u);
        F w = F::mul(u, v);
        F s = F::mul(ax, v);
        F xx = F::sqr(ax);
        F m = F::add(F::dbl(xx), xx);
        XYZZ r;
        r.x = F::sub(F::sub(F::sqr(m), s), s);
        r.y = F::sub(F::mul(m, F::sub(s, r.x)), F::mul(w, ay));
        r.zz = v;
        r.zzz = w;
        return r;
    }
    // dbl-2008-s-1, a = 0
    CZK_HD_NOINLINE static XYZZ dbl(const XYZZ& p) {
        if (p.is_inf()) return p;
        F u = F::dbl(p.y);
        F v = F::sqr(u);
        F w = F::mul(u, v);
        F s = F::mul(p.x, v);
        F xx = F::sqr(p.x);
        F m = F::add(F::dbl(xx), xx);
        XYZZ r;
        r.x = F::sub(F::sub(F::sqr(m), s), s);
        r.y = F::sub(F::mul(m, F::sub(s, r.x)), F::mul(w, p.y));
        r.zz = F::mul(v, p.zz);
        r.zzz = F::mul(w, p.zzz);
        return r;
    }
    // this += (ax, ay)   (madd-2008-s); (ax, ay) must be a finite point
    CZK_HD void add_affine(const F& ax, const F& ay) {
        if (is_inf()) {
            *this = from_affine(ax, ay);
            return;
        }
        F u2 = F::mul(ax, zz);
        F s2 = F::mul(ay, zzz);
        F p = F::sub(u2, x);
        F r = F::sub(s2, y);
        if (p.is_zero() && r.is_zero()) {
            *this = dbl_affine(ax, ay);
            return;
        }
        F pp = F::sqr(p);
        F ppp = F::mul(p, pp);
        F q = F::mul(x, pp);
        F x3 = F::sub(F::sub(F::sub(F::sqr(r), ppp), q), q);
        F y3 = F::sub(F::mul(r, F::sub(q, x3)), F::mul(y, ppp));
        x = x3;
        y = y3;
        zz = F::mul(zz, pp);
        zzz = F::mul(zzz, ppp);
    }
    // this += o   (add-2008-s)
    CZK_HD_NOINLINE void add(const XYZZ& o) {
        if (o.is_inf()) return;
        if (is_inf()) {
            *this = o;
            return;
        }
        F u1 = F::mul(x, o.zz);
        F u2 = F::mul(o.x, zz);
        F s1 = F::mul(y, o.zzz);
        F s2 = F::mul(o.y, zzz);
        F p = F::sub(u2, u1);
        F r = F::sub(s2, s1);
        if (p.is_zero() && r.is_zero()) {
            *this = dbl(*this);
            return;
        }
        F pp = F::sqr(p);
        F ppp = F::mul(p, pp);
        F q = F::mul(u1, pp);
        F x3 = F::sub(F::sub(F::sub(F::sqr(r), ppp), q), q);
        F y3 = F::sub(F::mul(r, F::sub(q, x3)), F::mul(s1, ppp));
        x = x3;
        y = y3;
        zz = F::mul(F::mul(zz, o.zz), pp);
        zzz = F::mul(F::mul(zzz, o.zzz), ppp);
    }
    CZK_HD void negate() { y = F::neg(y); }
};

}  // namespace czk
