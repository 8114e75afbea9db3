/* Prime-field "template": include with
 *     #define FP   <prefix>      (fr | fq)
 *     #define NL   <limbs>       (4 | 6)
 * TEST INFRASTRUCTURE (CPU oracle) - see czk_oracle.c for the rules.
 *
 * Restates, limb for limb (u64 limbs, u128 products), the reference's
 *   algebra/ff/src/fields/arithmetic.rs:7-57     mul_assign (no-carry CIOS branch; both moduli qualify)
 *   algebra/ff/src/fields/arithmetic.rs:59-82    into_repr  (Montgomery reduction)
 *   algebra/ff/src/fields/macros.rs:242-246      reduce
 *   algebra/ff/src/fields/macros.rs:298-304      double_in_place
 *   algebra/ff/src/fields/macros.rs:368-422      inverse (binary extended Euclid, Guajardo et al. alg. 16)
 *   algebra/ff/src/fields/macros.rs:444-454      from_repr
 *   algebra/ff/src/fields/macros.rs:605-618      neg
 *   algebra/ff/src/fields/macros.rs:663-682      add_assign / sub_assign
 * square_in_place (arithmetic.rs:84-171) is a different schedule of the same
 * product followed by the same reduction; a field element has one canonical
 * Montgomery encoding, so it is restated as mul(a, a).
 */
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(FP, name)

typedef struct {
    uint64_t l[NL];
} FN(t);

static uint64_t FN(MOD)[NL];
static uint64_t FN(INV);
static FN(t) FN(R);  /* Montgomery one */
static FN(t) FN(R2); /* R^2 mod p */

static inline int FN(is_zero)(const FN(t) * a) {
    uint64_t o = 0;
    for (int i = 0; i < NL; i++) o |= a->l[i];
    return o == 0;
}
static inline int FN(eq)(const FN(t) * a, const FN(t) * b) {
    uint64_t o = 0;
    for (int i = 0; i < NL; i++) o |= a->l[i] ^ b->l[i];
    return o == 0;
}
/* BigInteger::cmp, most significant limb first */
static inline int FN(cmp_raw)(const uint64_t *a, const uint64_t *b) {
    for (int i = NL - 1; i >= 0; i--) {
        if (a[i] < b[i]) return -1;
        if (a[i] > b[i]) return 1;
    }
    return 0;
}
static inline uint64_t FN(add_nocarry)(uint64_t *a, const uint64_t *b) {
    u128 c = 0;
    for (int i = 0; i < NL; i++) {
        c += (u128)a[i] + b[i];
        a[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
static inline uint64_t FN(sub_noborrow)(uint64_t *a, const uint64_t *b) {
    uint64_t borrow = 0;
    for (int i = 0; i < NL; i++) {
        u128 d = (u128)a[i] - b[i] - borrow;
        a[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}
static inline void FN(div2_raw)(uint64_t *a) {
    uint64_t t = 0;
    for (int i = NL - 1; i >= 0; i--) {
        uint64_t t2 = a[i] << 63;
        a[i] = (a[i] >> 1) | t;
        t = t2;
    }
}
static inline void FN(mul2_raw)(uint64_t *a) {
    uint64_t last = 0;
    for (int i = 0; i < NL; i++) {
        uint64_t tmp = a[i] >> 63;
        a[i] = (a[i] << 1) | last;
        last = tmp;
    }
}
/* macros.rs:242-246 */
static inline void FN(reduce)(FN(t) * a) {
    if (FN(cmp_raw)(a->l, FN(MOD)) >= 0) FN(sub_noborrow)(a->l, FN(MOD));
}
/* macros.rs:663-671 */
static inline void FN(add)(FN(t) * r, const FN(t) * a, const FN(t) * b) {
    FN(t) t = *a;
    FN(add_nocarry)(t.l, b->l);
    FN(reduce)(&t);
    *r = t;
}
/* macros.rs:673-682 */
static inline void FN(sub)(FN(t) * r, const FN(t) * a, const FN(t) * b) {
    FN(t) t = *a;
    if (FN(cmp_raw)(b->l, t.l) > 0) FN(add_nocarry)(t.l, FN(MOD));
    FN(sub_noborrow)(t.l, b->l);
    *r = t;
}
/* macros.rs:298-304 */
static inline void FN(dbl)(FN(t) * r, const FN(t) * a) {
    FN(t) t = *a;
    FN(mul2_raw)(t.l);
    FN(reduce)(&t);
    *r = t;
}
/* macros.rs:605-618 */
static inline void FN(neg)(FN(t) * r, const FN(t) * a) {
    if (!FN(is_zero)(a)) {
        FN(t) t;
        memcpy(t.l, FN(MOD), sizeof t.l);
        FN(sub_noborrow)(t.l, a->l);
        *r = t;
    } else {
        *r = *a;
    }
}
/* arithmetic.rs:36-52 (no-carry CIOS) */
static inline void FN(mul)(FN(t) * out, const FN(t) * a, const FN(t) * b) {
    uint64_t r[NL];
    for (int i = 0; i < NL; i++) r[i] = 0;
    for (int i = 0; i < NL; i++) {
        u128 t = (u128)r[0] + (u128)a->l[0] * b->l[i]; /* fa::mac */
        r[0] = (uint64_t)t;
        uint64_t carry1 = (uint64_t)(t >> 64);
        uint64_t k = r[0] * FN(INV);
        t = (u128)r[0] + (u128)k * FN(MOD)[0]; /* fa::mac_discard */
        uint64_t carry2 = (uint64_t)(t >> 64);
        for (int j = 1; j < NL; j++) {
            t = (u128)r[j] + (u128)a->l[j] * b->l[i] + carry1; /* mac_with_carry */
            r[j] = (uint64_t)t;
            carry1 = (uint64_t)(t >> 64);
            t = (u128)r[j] + (u128)k * FN(MOD)[j] + carry2;
            r[j - 1] = (uint64_t)t;
            carry2 = (uint64_t)(t >> 64);
        }
        r[NL - 1] = carry1 + carry2;
    }
    memcpy(out->l, r, sizeof r);
    FN(reduce)(out);
}
static inline void FN(sqr)(FN(t) * r, const FN(t) * a) { FN(mul)(r, a, a); }

/* arithmetic.rs:59-82: Montgomery form -> canonical integer */
static inline void FN(into_repr)(uint64_t *out, const FN(t) * a) {
    uint64_t r[NL];
    memcpy(r, a->l, sizeof r);
    for (int i = 0; i < NL; i++) {
        uint64_t k = r[i] * FN(INV);
        u128 t = (u128)r[i] + (u128)k * FN(MOD)[0];
        uint64_t carry = (uint64_t)(t >> 64);
        for (int j = 1; j < NL; j++) {
            int idx = (j + i) % NL;
            t = (u128)r[idx] + (u128)k * FN(MOD)[j] + carry;
            r[idx] = (uint64_t)t;
            carry = (uint64_t)(t >> 64);
        }
        r[i % NL] = carry;
    }
    memcpy(out, r, sizeof r);
}
/* macros.rs:444-454: canonical integer (< p) -> Montgomery form; returns 0 if not valid */
static inline int FN(from_repr)(FN(t) * out, const uint64_t *repr) {
    FN(t) r;
    memcpy(r.l, repr, sizeof r.l);
    if (FN(is_zero)(&r)) {
        *out = r;
        return 1;
    }
    if (FN(cmp_raw)(r.l, FN(MOD)) >= 0) return 0;
    FN(mul)(out, &r, &FN(R2));
    return 1;
}
/* macros.rs:368-422 */
static inline int FN(inv)(FN(t) * out, const FN(t) * a) {
    if (FN(is_zero)(a)) return 0;
    uint64_t one[NL] = {1};
    uint64_t u[NL], v[NL];
    memcpy(u, a->l, sizeof u);
    memcpy(v, FN(MOD), sizeof v);
    FN(t) b = FN(R2);
    FN(t) c;
    memset(&c, 0, sizeof c);
    while (FN(cmp_raw)(u, one) != 0 && FN(cmp_raw)(v, one) != 0) {
        while ((u[0] & 1) == 0) {
            FN(div2_raw)(u);
            if ((b.l[0] & 1) == 0) {
                FN(div2_raw)(b.l);
            } else {
                FN(add_nocarry)(b.l, FN(MOD));
                FN(div2_raw)(b.l);
            }
        }
        while ((v[0] & 1) == 0) {
            FN(div2_raw)(v);
            if ((c.l[0] & 1) == 0) {
                FN(div2_raw)(c.l);
            } else {
                FN(add_nocarry)(c.l, FN(MOD));
                FN(div2_raw)(c.l);
            }
        }
        if (FN(cmp_raw)(v, u) < 0) {
            FN(sub_noborrow)(u, v);
            FN(sub)(&b, &b, &c);
        } else {
            FN(sub_noborrow)(v, u);
            FN(sub)(&c, &c, &b);
        }
    }
    *out = (FN(cmp_raw)(u, one) == 0) ? b : c;
    return 1;
}
/* Field::pow over little-endian u64 exponent limbs (ff/src/fields/mod.rs pow: MSB-first square-and-multiply) */
static inline void FN(pow)(FN(t) * out, const FN(t) * a, const uint64_t *exp, int nexp) {
    FN(t) res = FN(R);
    int started = 0;
    for (int i = nexp * 64 - 1; i >= 0; i--) {
        int bit = (exp[i / 64] >> (i % 64)) & 1;
        if (started) FN(sqr)(&res, &res);
        if (bit) {
            started = 1;
            FN(mul)(&res, &res, a);
        }
    }
    *out = res;
}
static inline void FN(from_u64)(FN(t) * out, uint64_t x) {
    uint64_t repr[NL] = {x};
    FN(from_repr)(out, repr);
}
/* derive INV = -p^-1 mod 2^64, R = 2^(64 NL) mod p, R2 = R^2 mod p from the modulus alone
 * (the tests pin them against the literals in the reference's parameter files). */
static void FN(init)(const uint64_t *modulus) {
    memcpy(FN(MOD), modulus, sizeof(uint64_t) * NL);
    uint64_t inv = 1;
    for (int i = 0; i < 63; i++) {
        inv = inv * inv;
        inv = inv * modulus[0];
    }
    FN(INV) = (uint64_t)0 - inv;
    FN(t) x;
    memset(&x, 0, sizeof x);
    x.l[0] = 1;
    for (int i = 0; i < 64 * NL; i++) {
        FN(mul2_raw)(x.l);
        FN(reduce)(&x);
    }
    FN(R) = x;
    for (int i = 0; i < 64 * NL; i++) {
        FN(mul2_raw)(x.l);
        FN(reduce)(&x);
    }
    FN(R2) = x;
}
#undef FN
#undef CAT
#undef CAT_
