/* Short-Weierstrass (a = 0) Jacobian group "template": include with
 *     #define G    <prefix>   (g1 | g2)
 *     #define BF   <base field prefix> (fq | fq2)
 *     #define BFW  <u64 words per base-field element> (6 | 12)
 * TEST INFRASTRUCTURE (CPU oracle) - see czk_oracle.c for the rules.
 *
 * Restates algebra/ec/src/models/short_weierstrass_jacobian.rs:
 *   :440-458  zero() = (1,1,0), is_zero() <=> Z == 0
 *   :502-535  double_in_place, COEFF_A == 0 branch
 *   :570-638  add_assign_mixed  (madd-2007-bl, with the P==Q -> double branch)
 *   :666-729  add_assign        (add-2007-bl,  with the P==Q -> double branch)
 *   :768-790  From<Projective> for Affine
 * and algebra/ec/src/msm/variable_base.rs:12-106 (Pippenger, unsigned windows).
 */
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define GN(name) CAT(G, name)
#define F(name) CAT(BF, name)

typedef struct {
    F(t) x, y;
    int inf;
} GN(aff);
typedef struct {
    F(t) x, y, z;
} GN(jac);

static inline void GN(jac_zero)(GN(jac) * p) {
    p->x = F(one)();
    p->y = F(one)();
    memset(&p->z, 0, sizeof p->z);
}
static inline int GN(jac_is_zero)(const GN(jac) * p) { return F(is_zero)(&p->z); }

/* :502-535 */
static inline void GN(jac_double)(GN(jac) * p) {
    if (GN(jac_is_zero)(p)) return;
    F(t) a, b, c, d, e, f, t;
    F(sqr)(&a, &p->x);      /* A = X1^2 */
    F(sqr)(&b, &p->y);      /* B = Y1^2 */
    F(sqr)(&c, &b);         /* C = B^2 */
    F(add)(&t, &p->x, &b);  /* D = 2*((X1+B)^2-A-C) */
    F(sqr)(&t, &t);
    F(sub)(&t, &t, &a);
    F(sub)(&t, &t, &c);
    F(dbl)(&d, &t);
    F(dbl)(&t, &a);         /* E = 3*A */
    F(add)(&e, &a, &t);
    F(sqr)(&f, &e);         /* F = E^2 */
    F(mul)(&p->z, &p->z, &p->y); /* Z3 = 2*Y1*Z1 */
    F(dbl)(&p->z, &p->z);
    F(sub)(&t, &f, &d);     /* X3 = F-2*D */
    F(sub)(&p->x, &t, &d);
    F(sub)(&t, &d, &p->x);  /* Y3 = E*(D-X3)-8*C */
    F(mul)(&t, &t, &e);
    F(dbl)(&c, &c);
    F(dbl)(&c, &c);
    F(dbl)(&c, &c);
    F(sub)(&p->y, &t, &c);
}

/* :570-638 */
static inline void GN(jac_add_mixed)(GN(jac) * p, const GN(aff) * o) {
    if (o->inf) return;
    if (GN(jac_is_zero)(p)) {
        p->x = o->x;
        p->y = o->y;
        p->z = F(one)();
        return;
    }
    F(t) z1z1, u2, s2;
    F(sqr)(&z1z1, &p->z);
    F(mul)(&u2, &o->x, &z1z1);
    F(mul)(&s2, &o->y, &p->z);
    F(mul)(&s2, &s2, &z1z1);
    if (F(eq)(&p->x, &u2) && F(eq)(&p->y, &s2)) {
        GN(jac_double)(p);
        return;
    }
    F(t) h, hh, i, j, r, v;
    F(sub)(&h, &u2, &p->x);
    F(sqr)(&hh, &h);
    F(dbl)(&i, &hh);
    F(dbl)(&i, &i);
    F(mul)(&j, &h, &i);
    F(sub)(&r, &s2, &p->y);
    F(dbl)(&r, &r);
    F(mul)(&v, &p->x, &i);
    F(sqr)(&p->x, &r);
    F(sub)(&p->x, &p->x, &j);
    F(sub)(&p->x, &p->x, &v);
    F(sub)(&p->x, &p->x, &v);
    F(mul)(&j, &j, &p->y);
    F(dbl)(&j, &j);
    F(sub)(&p->y, &v, &p->x);
    F(mul)(&p->y, &p->y, &r);
    F(sub)(&p->y, &p->y, &j);
    F(add)(&p->z, &p->z, &h);
    F(sqr)(&p->z, &p->z);
    F(sub)(&p->z, &p->z, &z1z1);
    F(sub)(&p->z, &p->z, &hh);
}

/* :666-729 */
static inline void GN(jac_add)(GN(jac) * p, const GN(jac) * o) {
    if (GN(jac_is_zero)(p)) {
        *p = *o;
        return;
    }
    if (GN(jac_is_zero)(o)) return;
    F(t) z1z1, z2z2, u1, u2, s1, s2;
    F(sqr)(&z1z1, &p->z);
    F(sqr)(&z2z2, &o->z);
    F(mul)(&u1, &p->x, &z2z2);
    F(mul)(&u2, &o->x, &z1z1);
    F(mul)(&s1, &p->y, &o->z);
    F(mul)(&s1, &s1, &z2z2);
    F(mul)(&s2, &o->y, &p->z);
    F(mul)(&s2, &s2, &z1z1);
    if (F(eq)(&u1, &u2) && F(eq)(&s1, &s2)) {
        GN(jac_double)(p);
        return;
    }
    F(t) h, i, j, r, v, t;
    F(sub)(&h, &u2, &u1);
    F(dbl)(&i, &h);
    F(sqr)(&i, &i);
    F(mul)(&j, &h, &i);
    F(sub)(&r, &s2, &s1);
    F(dbl)(&r, &r);
    F(mul)(&v, &u1, &i);
    F(sqr)(&t, &r);
    F(sub)(&t, &t, &j);
    F(t) v2;
    F(dbl)(&v2, &v);
    F(sub)(&p->x, &t, &v2);
    F(sub)(&t, &v, &p->x);
    F(mul)(&t, &r, &t);
    F(mul)(&s1, &s1, &j);
    F(dbl)(&s1, &s1);
    F(sub)(&p->y, &t, &s1);
    F(add)(&t, &p->z, &o->z);
    F(sqr)(&t, &t);
    F(sub)(&t, &t, &z1z1);
    F(sub)(&t, &t, &z2z2);
    F(mul)(&p->z, &t, &h);
}

static inline void GN(jac_neg)(GN(jac) * p) {
    if (!GN(jac_is_zero)(p)) F(neg)(&p->y, &p->y);
}

/* :147-157 affine zero = (0, 1, inf) ; :768-790 */
static inline void GN(to_affine)(GN(aff) * a, const GN(jac) * p) {
    if (GN(jac_is_zero)(p)) {
        memset(&a->x, 0, sizeof a->x);
        a->y = F(one)();
        a->inf = 1;
        return;
    }
    F(t) one = F(one)();
    if (F(eq)(&p->z, &one)) {
        a->x = p->x;
        a->y = p->y;
        a->inf = 0;
        return;
    }
    F(t) zinv, zinv2, zinv3;
    F(inv)(&zinv, &p->z);
    F(sqr)(&zinv2, &zinv);
    F(mul)(&a->x, &p->x, &zinv2);
    F(mul)(&zinv3, &zinv2, &zinv);
    F(mul)(&a->y, &p->y, &zinv3);
    a->inf = 0;
}
static inline void GN(from_affine)(GN(jac) * p, const GN(aff) * a) {
    if (a->inf) {
        GN(jac_zero)(p);
    } else {
        p->x = a->x;
        p->y = a->y;
        p->z = F(one)();
    }
}

/* ProjectiveCurve::mul / AffineCurve::mul: MSB-first double-and-add over the
 * canonical scalar bits (ec/src/models/short_weierstrass_jacobian.rs mul_bits). */
static void GN(scalar_mul)(GN(jac) * out, const GN(aff) * base, const uint64_t repr[4]) {
    GN(jac) res;
    GN(jac_zero)(&res);
    for (int i = 255; i >= 0; i--) {
        GN(jac_double)(&res);
        if ((repr[i / 64] >> (i % 64)) & 1) GN(jac_add_mixed)(&res, base);
    }
    *out = res;
}
static void GN(jac_scalar_mul)(GN(jac) * out, const GN(jac) * base, const uint64_t repr[4]) {
    GN(jac) res;
    GN(jac_zero)(&res);
    for (int i = 255; i >= 0; i--) {
        GN(jac_double)(&res);
        if ((repr[i / 64] >> (i % 64)) & 1) GN(jac_add)(&res, base);
    }
    *out = res;
}

/* external array layout: point i = BFW words x | BFW words y (Montgomery), inf[i] byte (may be NULL) */
static inline void GN(load_aff)(GN(aff) * a, const uint64_t *xy, const uint8_t *inf, size_t i) {
    memcpy(&a->x, xy + 2 * BFW * i, 8 * BFW);
    memcpy(&a->y, xy + 2 * BFW * i + BFW, 8 * BFW);
    a->inf = inf ? inf[i] : 0;
}
static inline void GN(store_aff)(uint64_t *xy, uint8_t *inf, size_t i, const GN(aff) * a) {
    memcpy(xy + 2 * BFW * i, &a->x, 8 * BFW);
    memcpy(xy + 2 * BFW * i + BFW, &a->y, 8 * BFW);
    if (inf) inf[i] = (uint8_t)a->inf;
}

/* variable_base.rs:12-106.  `scalars` are canonical BigInt256 (4 words each).
 * threads > 1 parallelises over windows exactly where the reference's dormant
 * Rayon path does (cfg_into_iter!(window_starts), :36). */
static void GN(msm_bigint)(GN(jac) * out, const uint64_t *bases_xy, const uint8_t *inf, const uint64_t *scalars,
                           size_t size, int threads) {
    size_t c = size < 32 ? 3 : (size_t)ark_ln_without_floats(size) + 2;
    const int num_bits = 253;
    uint64_t fr_one[4] = {1, 0, 0, 0};
    int nwin = (int)((num_bits + c - 1) / c);
    GN(jac) *window_sums = (GN(jac) *)malloc(sizeof(GN(jac)) * (size_t)nwin);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1)
    for (int w = 0; w < nwin; w++) {
        size_t w_start = (size_t)w * c;
        GN(jac) res;
        GN(jac_zero)(&res);
        size_t nb = ((size_t)1 << c) - 1;
        GN(jac) *buckets = (GN(jac) *)malloc(sizeof(GN(jac)) * nb);
        for (size_t b = 0; b < nb; b++) GN(jac_zero)(&buckets[b]);
        for (size_t i = 0; i < size; i++) {
            const uint64_t *s = scalars + 4 * i;
            if ((s[0] | s[1] | s[2] | s[3]) == 0) continue; /* .filter(|(s, _)| !s.is_zero()) */
            GN(aff) base;
            GN(load_aff)(&base, bases_xy, inf, i);
            if (s[0] == fr_one[0] && s[1] == 0 && s[2] == 0 && s[3] == 0) {
                if (w_start == 0) GN(jac_add_mixed)(&res, &base);
            } else {
                /* scalar.divn(w_start); scalar.as_ref()[0] % (1 << c) */
                size_t limb = w_start / 64, sh = w_start % 64;
                uint64_t lo = s[limb] >> sh;
                if (sh && limb + 1 < 4) lo |= s[limb + 1] << (64 - sh);
                uint64_t digit = lo % ((uint64_t)1 << c);
                if (digit != 0) GN(jac_add_mixed)(&buckets[digit - 1], &base);
            }
        }
        GN(jac) running;
        GN(jac_zero)(&running);
        for (size_t b = nb; b-- > 0;) {
            GN(jac_add)(&running, &buckets[b]);
            GN(jac_add)(&res, &running);
        }
        free(buckets);
        window_sums[w] = res;
    }
    GN(jac) total;
    GN(jac_zero)(&total);
    for (int w = nwin - 1; w >= 1; w--) {
        GN(jac_add)(&total, &window_sums[w]);
        for (size_t k = 0; k < c; k++) GN(jac_double)(&total);
    }
    GN(jac) lowest = window_sums[0];
    GN(jac_add)(&lowest, &total);
    free(window_sums);
    *out = lowest;
}

#undef GN
#undef F
#undef CAT
#undef CAT_
