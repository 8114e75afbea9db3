/* czk_oracle.c - CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and there only as the checker / the CPU baseline.
 * The product (collaborative-zksnark_b200/) never links or calls it.
 *
 * The reference (alex-ozdemir/collaborative-zksnark @ 8cff2c2) is pure Rust and
 * cannot be compiled in the build image (no rustc/cargo, nightly-only features,
 * crates.io dependencies not vendored), so this file restates its in-tree
 * algorithms in plain C (u64 limbs, unsigned __int128 products).  Each function
 * cites the reference file:line it follows (paths relative to /root/reference).
 *
 * Pinning: the reference holds no golden vectors for MSM / NTT / proofs
 * (SURVEY.md section 8c).  What it does hold - the Montgomery-form parameter
 * literals in curves/bls12_377/src/ - is extracted into
 * tests/golden/bls12_377_constants.json and checked against this file's
 * self-derived constants; the algorithms are pinned by the identities the
 * reference's own tests use (MSM == naive sum, FFT == Horner evaluation,
 * FFT == serial CLRS radix-2, Beaver product correctness) and by an independent
 * Python big-int model (oracle/pymodel.py).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
#define EXPORT __attribute__((visibility("default")))

/* utils/src/lib.rs:65-73 (ark_std::log2: ceil(log2 x)) and ec/src/msm/mod.rs:10-13 */
static uint32_t ark_log2(size_t x) {
    if (x == 0) return 0;
    if ((x & (x - 1)) == 0) return (uint32_t)__builtin_ctzll(x);
    return 64u - (uint32_t)__builtin_clzll(x);
}
static size_t ark_ln_without_floats(size_t a) { return (size_t)(ark_log2(a) * 69 / 100); }

/* ---------------------------------------------------------------- fields */
#define FP fr
#define NL 4
#include "fp_tmpl.h"
#undef FP
#undef NL
#define FP fq
#define NL 6
#include "fp_tmpl.h"
#undef FP
#undef NL

/* curves/bls12_377/src/fields/fr.rs:31-38, fq.rs:24-33 (MODULUS) */
static const uint64_t FR_MODULUS[4] = {725501752471715841ull, 6461107452199829505ull, 6968279316240510977ull,
                                       1345280370688173398ull};
static const uint64_t FQ_MODULUS[6] = {0x8508c00000000001ull, 0x170b5d4430000000ull, 0x1ef3622fba094800ull,
                                       0x1a22d9f300f5138full, 0xc63b05c06ca1493bull, 0x1ae3a4617c510eaull};
/* fr.rs:23-28 LARGE_SUBGROUP_ROOT_OF_UNITY (Montgomery form), SMALL_SUBGROUP_BASE = 3, adicity 1, TWO_ADICITY = 47 */
static const uint64_t FR_LARGE_SUBGROUP_ROOT[4] = {0x9bfe9d90c790c167ull, 0x7175a69e39013bffull, 0x3fbbb698adabcf93ull,
                                                   0xc59f8d8d6f0dc97ull};
#define FR_TWO_ADICITY 47
#define FR_SMALL_SUBGROUP_BASE 3
#define FR_SMALL_SUBGROUP_BASE_ADICITY 1
#define FR_GENERATOR_U64 22 /* fr.rs:65-74 decodes to 22 (the doc comment saying 11 is stale) */

static inline fq_t fq_one(void) { return fq_R; }
static inline fr_t fr_one(void) { return fr_R; }

/* Fq2 = Fq[u]/(u^2+5).  curves/bls12_377/src/fields/fq2.rs:13,29-34 ;
 * algebra/ff/src/fields/models/quadratic_extension.rs:569-583 (mul), :257-306 (square), :308-324 (inverse) */
typedef struct {
    fq_t c0, c1;
} fq2_t;
static inline void fq_mul_by_nonresidue(fq_t *r, const fq_t *fe) {
    /* let mut fe = -fe.double(); fe.double_in_place(); fe - original */
    fq_t t;
    fq_dbl(&t, fe);
    fq_neg(&t, &t);
    fq_dbl(&t, &t);
    fq_sub(r, &t, fe);
}
static inline fq2_t fq2_one(void) {
    fq2_t o;
    o.c0 = fq_R;
    memset(&o.c1, 0, sizeof o.c1);
    return o;
}
static inline int fq2_is_zero(const fq2_t *a) { return fq_is_zero(&a->c0) && fq_is_zero(&a->c1); }
static inline int fq2_eq(const fq2_t *a, const fq2_t *b) { return fq_eq(&a->c0, &b->c0) && fq_eq(&a->c1, &b->c1); }
static inline void fq2_add(fq2_t *r, const fq2_t *a, const fq2_t *b) {
    fq_add(&r->c0, &a->c0, &b->c0);
    fq_add(&r->c1, &a->c1, &b->c1);
}
static inline void fq2_sub(fq2_t *r, const fq2_t *a, const fq2_t *b) {
    fq_sub(&r->c0, &a->c0, &b->c0);
    fq_sub(&r->c1, &a->c1, &b->c1);
}
static inline void fq2_dbl(fq2_t *r, const fq2_t *a) {
    fq_dbl(&r->c0, &a->c0);
    fq_dbl(&r->c1, &a->c1);
}
static inline void fq2_neg(fq2_t *r, const fq2_t *a) {
    fq_neg(&r->c0, &a->c0);
    fq_neg(&r->c1, &a->c1);
}
static inline void fq2_mul(fq2_t *r, const fq2_t *a, const fq2_t *b) {
    fq_t v0, v1, t, u, nr;
    fq_mul(&v0, &a->c0, &b->c0);
    fq_mul(&v1, &a->c1, &b->c1);
    fq_add(&t, &a->c1, &a->c0);
    fq_add(&u, &b->c0, &b->c1);
    fq_mul(&t, &t, &u);
    fq_sub(&t, &t, &v0);
    fq_sub(&t, &t, &v1);
    fq_mul_by_nonresidue(&nr, &v1);
    fq_add(&r->c0, &v0, &nr);
    r->c1 = t;
}
static inline void fq2_sqr(fq2_t *r, const fq2_t *a) {
    /* generic-beta branch */
    fq_t v0, v3, v2, nr, t;
    fq_sub(&v0, &a->c0, &a->c1);
    fq_mul_by_nonresidue(&nr, &a->c1);
    fq_sub(&v3, &a->c0, &nr);
    fq_mul(&v2, &a->c0, &a->c1);
    fq_mul(&v0, &v0, &v3);
    fq_dbl(&r->c1, &v2);
    /* add_and_mul_base_field_by_nonresidue_plus_one(v0, v2) = (v0 + v2) + beta*v2 */
    fq_add(&t, &v0, &v2);
    fq_mul_by_nonresidue(&nr, &v2);
    fq_add(&r->c0, &t, &nr);
}
static inline int fq2_inv(fq2_t *r, const fq2_t *a) {
    if (fq2_is_zero(a)) return 0;
    fq_t v1, v0, nr, inv;
    fq_sqr(&v1, &a->c1);
    fq_sqr(&v0, &a->c0);
    fq_mul_by_nonresidue(&nr, &v1);
    fq_sub(&v0, &v0, &nr);
    if (!fq_inv(&inv, &v0)) return 0;
    fq_mul(&r->c0, &a->c0, &inv);
    fq_mul(&v1, &a->c1, &inv);
    fq_neg(&r->c1, &v1);
    return 1;
}

/* ---------------------------------------------------------------- groups */
#define G g1
#define BF fq
#define BFW 6
#include "ec_tmpl.h"
#undef G
#undef BF
#undef BFW
#define G g2
#define BF fq2
#define BFW 12
#include "ec_tmpl.h"
#undef G
#undef BF
#undef BFW


/* curves/bls12_377/src/curves/g1.rs:46-51 and g2.rs:68-86: generators, canonical integers (converted to
 * Montgomery form at init) */
static const uint64_t G1_GEN_CANON[12] = {0xeab9b16eb21be9efull, 0xd5481512ffcd394eull, 0x188282c8bd37cb5cull, 0x85951e2caa9d41bbull, 0xc8fc6225bf87ff54ull, 0x008848defe740a67ull,
                                          0xfd82de55559c8ea6ull, 0xc2fe3d3634a9591aull, 0x6d182ad44fb82305ull, 0xbd7fb348ca3e52d9ull, 0x1f674f5d30afeec4ull, 0x01914a69c5102effull};
static const uint64_t G2_GEN_CANON[24] = {0x74e3e48f7c005196ull, 0x71889f52bb535402ull, 0x7ea501f557db6b9bull, 0xc565f071203e5031ull, 0xc89630a2a3841d01ull, 0x018480be71c785feull,
                                          0xb26bfefa6ea16afeull, 0x5cf89984bff76fe6ull, 0xe7223ece0799c9deull, 0x532777ee6651cecbull, 0x70dc5a51b1b140d5ull, 0x00ea6040e7004031ull,
                                          0xf094094409fd4ddfull, 0xf2cf88886d8c7c2eull, 0xe458c282f832d204ull, 0xde03ed7274b49a58ull, 0xd960736bcbb2efb4ull, 0x00690d665d446f7bull,
                                          0xd9a1cdd185eb8f93ull, 0x4279b83f5e52270bull, 0x2463b01acee304c2ull, 0x61ef11ac3d591bf1ull, 0x9e549da3151a70aaull, 0x00f8169fd2835518ull};
uint64_t ORC_G1_GEN[12], ORC_G2_GEN[24];
static void orc_init_generators(void) {
    for (int i = 0; i < 2; i++) fq_from_repr((fq_t *)(ORC_G1_GEN + 6 * i), G1_GEN_CANON + 6 * i);
    for (int i = 0; i < 4; i++) fq_from_repr((fq_t *)(ORC_G2_GEN + 6 * i), G2_GEN_CANON + 6 * i);
}
EXPORT void orc_generators(uint64_t *g1, uint64_t *g2) {
    memcpy(g1, ORC_G1_GEN, sizeof ORC_G1_GEN);
    memcpy(g2, ORC_G2_GEN, sizeof ORC_G2_GEN);
}

static int g_inited = 0;
EXPORT void orc_init(void) {
    if (g_inited) return;
    fr_init(FR_MODULUS);
    fq_init(FQ_MODULUS);
    orc_init_generators();
    g_inited = 1;
}
__attribute__((constructor)) static void orc_ctor(void) { orc_init(); }

/* derived parameter readout, for the golden-constant tests */
EXPORT void orc_params(uint64_t *fr_R_out, uint64_t *fr_R2_out, uint64_t *fr_inv, uint64_t *fq_R_out,
                       uint64_t *fq_R2_out, uint64_t *fq_inv) {
    memcpy(fr_R_out, fr_R.l, 32);
    memcpy(fr_R2_out, fr_R2.l, 32);
    *fr_inv = fr_INV;
    memcpy(fq_R_out, fq_R.l, 48);
    memcpy(fq_R2_out, fq_R2.l, 48);
    *fq_inv = fq_INV;
}

/* ------------------------------------------------ elementwise field API (arrays of Montgomery limbs) */
EXPORT void orc_fr_mul(uint64_t *r, const uint64_t *a, const uint64_t *b, size_t n) {
    for (size_t i = 0; i < n; i++) fr_mul((fr_t *)(r + 4 * i), (const fr_t *)(a + 4 * i), (const fr_t *)(b + 4 * i));
}
EXPORT void orc_fr_add(uint64_t *r, const uint64_t *a, const uint64_t *b, size_t n) {
    for (size_t i = 0; i < n; i++) fr_add((fr_t *)(r + 4 * i), (const fr_t *)(a + 4 * i), (const fr_t *)(b + 4 * i));
}
EXPORT void orc_fr_sub(uint64_t *r, const uint64_t *a, const uint64_t *b, size_t n) {
    for (size_t i = 0; i < n; i++) fr_sub((fr_t *)(r + 4 * i), (const fr_t *)(a + 4 * i), (const fr_t *)(b + 4 * i));
}
EXPORT void orc_fr_neg(uint64_t *r, const uint64_t *a, size_t n) {
    for (size_t i = 0; i < n; i++) fr_neg((fr_t *)(r + 4 * i), (const fr_t *)(a + 4 * i));
}
EXPORT int orc_fr_inv(uint64_t *r, const uint64_t *a, size_t n) {
    int ok = 1;
    for (size_t i = 0; i < n; i++) ok &= fr_inv((fr_t *)(r + 4 * i), (const fr_t *)(a + 4 * i));
    return ok;
}
EXPORT int orc_fr_from_repr(uint64_t *r, const uint64_t *a, size_t n) {
    int ok = 1;
    for (size_t i = 0; i < n; i++) ok &= fr_from_repr((fr_t *)(r + 4 * i), a + 4 * i);
    return ok;
}
EXPORT void orc_fr_into_repr(uint64_t *r, const uint64_t *a, size_t n) {
    for (size_t i = 0; i < n; i++) fr_into_repr(r + 4 * i, (const fr_t *)(a + 4 * i));
}
EXPORT void orc_fq_mul(uint64_t *r, const uint64_t *a, const uint64_t *b, size_t n) {
    for (size_t i = 0; i < n; i++) fq_mul((fq_t *)(r + 6 * i), (const fq_t *)(a + 6 * i), (const fq_t *)(b + 6 * i));
}
EXPORT void orc_fq_add(uint64_t *r, const uint64_t *a, const uint64_t *b, size_t n) {
    for (size_t i = 0; i < n; i++) fq_add((fq_t *)(r + 6 * i), (const fq_t *)(a + 6 * i), (const fq_t *)(b + 6 * i));
}
EXPORT void orc_fq_sub(uint64_t *r, const uint64_t *a, const uint64_t *b, size_t n) {
    for (size_t i = 0; i < n; i++) fq_sub((fq_t *)(r + 6 * i), (const fq_t *)(a + 6 * i), (const fq_t *)(b + 6 * i));
}
EXPORT int orc_fq_inv(uint64_t *r, const uint64_t *a, size_t n) {
    int ok = 1;
    for (size_t i = 0; i < n; i++) ok &= fq_inv((fq_t *)(r + 6 * i), (const fq_t *)(a + 6 * i));
    return ok;
}
EXPORT int orc_fq_from_repr(uint64_t *r, const uint64_t *a, size_t n) {
    int ok = 1;
    for (size_t i = 0; i < n; i++) ok &= fq_from_repr((fq_t *)(r + 6 * i), a + 6 * i);
    return ok;
}
EXPORT void orc_fq_into_repr(uint64_t *r, const uint64_t *a, size_t n) {
    for (size_t i = 0; i < n; i++) fq_into_repr(r + 6 * i, (const fq_t *)(a + 6 * i));
}
EXPORT void orc_fq2_mul(uint64_t *r, const uint64_t *a, const uint64_t *b, size_t n) {
    for (size_t i = 0; i < n; i++) fq2_mul((fq2_t *)(r + 12 * i), (const fq2_t *)(a + 12 * i), (const fq2_t *)(b + 12 * i));
}
EXPORT void orc_fq2_sqr(uint64_t *r, const uint64_t *a, size_t n) {
    for (size_t i = 0; i < n; i++) fq2_sqr((fq2_t *)(r + 12 * i), (const fq2_t *)(a + 12 * i));
}
EXPORT int orc_fq2_inv(uint64_t *r, const uint64_t *a, size_t n) {
    int ok = 1;
    for (size_t i = 0; i < n; i++) ok &= fq2_inv((fq2_t *)(r + 12 * i), (const fq2_t *)(a + 12 * i));
    return ok;
}

/* ------------------------------------------------ group API.
 * Affine arrays: point i = x | y (Montgomery limbs; 12 words G1, 24 words G2) + inf[i] byte.
 * Jacobian points: x | y | z (18 / 36 words). */
#define GROUP_API(G, W)                                                                                              \
    EXPORT void orc_##G##_jac_double(uint64_t *p) { G##_jac_double((G##_jac *)p); }                                  \
    EXPORT void orc_##G##_jac_add(uint64_t *p, const uint64_t *o) { G##_jac_add((G##_jac *)p, (const G##_jac *)o); } \
    EXPORT void orc_##G##_jac_add_mixed(uint64_t *p, const uint64_t *xy, int inf) {                                  \
        G##_aff a;                                                                                                   \
        uint8_t f = (uint8_t)inf;                                                                                    \
        G##_load_aff(&a, xy, &f, 0);                                                                                 \
        G##_jac_add_mixed((G##_jac *)p, &a);                                                                         \
    }                                                                                                                \
    EXPORT int orc_##G##_jac_to_affine(uint64_t *xy, const uint64_t *p) {                                            \
        G##_aff a;                                                                                                   \
        G##_to_affine(&a, (const G##_jac *)p);                                                                       \
        G##_store_aff(xy, NULL, 0, &a);                                                                              \
        return a.inf;                                                                                                \
    }                                                                                                                \
    /* out = scalar * base, affine in / affine out; scalar is a Montgomery-form Fr */                                \
    EXPORT int orc_##G##_scalar_mul(uint64_t *out_xy, const uint64_t *base_xy, int base_inf, const uint64_t *s) {    \
        G##_aff a, o;                                                                                                \
        uint8_t f = (uint8_t)base_inf;                                                                               \
        uint64_t repr[4];                                                                                            \
        G##_jac r;                                                                                                   \
        G##_load_aff(&a, base_xy, &f, 0);                                                                            \
        fr_into_repr(repr, (const fr_t *)s);                                                                         \
        G##_scalar_mul(&r, &a, repr);                                                                                \
        G##_to_affine(&o, &r);                                                                                       \
        G##_store_aff(out_xy, NULL, 0, &o);                                                                          \
        return o.inf;                                                                                                \
    }                                                                                                                \
    /* AffineCurve::multi_scalar_mul (ec/src/lib.rs:302-311): Fr -> into_repr, then VariableBaseMSM; result as      \
     * affine (share/msm.rs:35 `.into()`).  scalars_montgomery = 0 means `scalars` already are BigInt256. */         \
    EXPORT int orc_##G##_msm(uint64_t *out_xy, const uint64_t *bases_xy, const uint8_t *inf, const uint64_t *scalars, \
                             int scalars_montgomery, size_t n, int threads) {                                        \
        uint64_t *repr = (uint64_t *)malloc(32 * (n ? n : 1));                                                       \
        if (scalars_montgomery) {                                                                                    \
            _Pragma("omp parallel for num_threads(threads) if (threads > 1)") for (size_t i = 0; i < n; i++)         \
                fr_into_repr(repr + 4 * i, (const fr_t *)(scalars + 4 * i));                                         \
        } else {                                                                                                     \
            memcpy(repr, scalars, 32 * n);                                                                           \
        }                                                                                                            \
        G##_jac r;                                                                                                   \
        G##_msm_bigint(&r, bases_xy, inf, repr, n, threads);                                                         \
        free(repr);                                                                                                  \
        G##_aff o;                                                                                                   \
        G##_to_affine(&o, &r);                                                                                       \
        G##_store_aff(out_xy, NULL, 0, &o);                                                                          \
        return o.inf;                                                                                                \
    }                                                                                                                \
    /* naive sum s_i P_i (algebra/test-templates/src/msm.rs:6-14) */                                                 \
    EXPORT int orc_##G##_msm_naive(uint64_t *out_xy, const uint64_t *bases_xy, const uint8_t *inf,                   \
                                   const uint64_t *scalars_mont, size_t n) {                                         \
        G##_jac acc;                                                                                                 \
        G##_jac_zero(&acc);                                                                                          \
        for (size_t i = 0; i < n; i++) {                                                                             \
            G##_aff a;                                                                                               \
            uint64_t repr[4];                                                                                        \
            G##_jac t;                                                                                               \
            G##_load_aff(&a, bases_xy, inf, i);                                                                      \
            fr_into_repr(repr, (const fr_t *)(scalars_mont + 4 * i));                                                \
            G##_scalar_mul(&t, &a, repr);                                                                            \
            G##_jac_add(&acc, &t);                                                                                   \
        }                                                                                                            \
        G##_aff o;                                                                                                   \
        G##_to_affine(&o, &acc);                                                                                     \
        G##_store_aff(out_xy, NULL, 0, &o);                                                                          \
        return o.inf;                                                                                                \
    }                                                                                                                \
    /* test-input generator (not a reference function): P_i = (k0 + i*kstep) * base, written as affine.              \
     * Distinct points of the prime-order subgroup when base is; batch-free (one inversion per point). */            \
    EXPORT void orc_##G##_gen_progression(uint64_t *out_xy, const uint64_t *base_xy, const uint64_t *k0_mont,        \
                                          const uint64_t *kstep_mont, size_t n, int threads) {                       \
        G##_aff base;                                                                                                \
        G##_load_aff(&base, base_xy, NULL, 0);                                                                       \
        uint64_t rs[4];                                                                                              \
        fr_into_repr(rs, (const fr_t *)kstep_mont);                                                                  \
        G##_jac stepj;                                                                                               \
        G##_scalar_mul(&stepj, &base, rs);                                                                           \
        G##_aff step;                                                                                                \
        G##_to_affine(&step, &stepj);                                                                                \
        size_t chunk = 4096;                                                                                         \
        size_t nchunks = (n + chunk - 1) / chunk;                                                                    \
        _Pragma("omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1)") for (size_t c = 0;    \
                                                                                                    c < nchunks;     \
                                                                                                    c++) {           \
            size_t lo = c * chunk, hi = lo + chunk < n ? lo + chunk : n;                                             \
            fr_t k, idx, t;                                                                                          \
            fr_from_u64(&idx, (uint64_t)lo);                                                                         \
            fr_mul(&t, &idx, (const fr_t *)kstep_mont);                                                              \
            fr_add(&k, &t, (const fr_t *)k0_mont);                                                                   \
            uint64_t r0[4];                                                                                          \
            fr_into_repr(r0, &k);                                                                                    \
            G##_jac cur;                                                                                             \
            G##_scalar_mul(&cur, &base, r0);                                                                         \
            for (size_t i = lo; i < hi; i++) {                                                                       \
                G##_aff o;                                                                                           \
                G##_to_affine(&o, &cur);                                                                             \
                if (o.inf) { /* encode infinity as (0,1) like the reference; callers also get no flag here */       \
                }                                                                                                    \
                G##_store_aff(out_xy, NULL, i, &o);                                                                  \
                G##_jac_add_mixed(&cur, &step);                                                                      \
            }                                                                                                        \
        }                                                                                                            \
    }

GROUP_API(g1, 6)
GROUP_API(g2, 12)

/* ------------------------------------------------ radix-2 domain + NTT */
typedef struct {
    uint64_t size;
    uint32_t log_size;
    fr_t size_as_fe, size_inv, group_gen, group_gen_inv, generator_inv, generator;
} domain_t;

static uint32_t k_adicity(size_t k, size_t n) {
    uint32_t r = 0;
    while (n > 1) {
        if (n % k == 0) {
            r++;
            n /= k;
        } else {
            return r;
        }
    }
    return r;
}
/* algebra/ff/src/fields/mod.rs:337-367 (large-subgroup branch; Fr sets SMALL_SUBGROUP_BASE) */
static int fr_get_root_of_unity(fr_t *out, size_t n) {
    size_t q = FR_SMALL_SUBGROUP_BASE;
    uint32_t q_adicity = k_adicity(q, n);
    size_t q_part = 1;
    for (uint32_t i = 0; i < q_adicity; i++) q_part *= q;
    uint32_t two_adicity = k_adicity(2, n);
    size_t two_part = (size_t)1 << two_adicity;
    if (n != two_part * q_part || two_adicity > FR_TWO_ADICITY || q_adicity > FR_SMALL_SUBGROUP_BASE_ADICITY) return 0;
    fr_t omega;
    memcpy(omega.l, FR_LARGE_SUBGROUP_ROOT, 32);
    for (uint32_t i = q_adicity; i < FR_SMALL_SUBGROUP_BASE_ADICITY; i++) {
        uint64_t e[1] = {q};
        fr_pow(&omega, &omega, e, 1);
    }
    for (uint32_t i = two_adicity; i < FR_TWO_ADICITY; i++) fr_sqr(&omega, &omega);
    *out = omega;
    return 1;
}
/* algebra/poly/src/domain/radix2/mod.rs:51-82 */
static int domain_new(domain_t *d, size_t num_coeffs) {
    size_t size = 1;
    while (size < num_coeffs) size <<= 1;
    d->size = size;
    d->log_size = (uint32_t)__builtin_ctzll(size);
    if (d->log_size > FR_TWO_ADICITY) return 0;
    if (!fr_get_root_of_unity(&d->group_gen, size)) return 0;
    fr_from_u64(&d->size_as_fe, size);
    fr_inv(&d->size_inv, &d->size_as_fe);
    fr_inv(&d->group_gen_inv, &d->group_gen);
    fr_from_u64(&d->generator, FR_GENERATOR_U64);
    fr_inv(&d->generator_inv, &d->generator);
    return 1;
}
EXPORT int orc_domain_params(size_t num_coeffs, uint64_t *size, uint64_t *group_gen, uint64_t *group_gen_inv,
                             uint64_t *size_inv, uint64_t *generator_inv) {
    domain_t d;
    if (!domain_new(&d, num_coeffs)) return 0;
    *size = d.size;
    memcpy(group_gen, d.group_gen.l, 32);
    memcpy(group_gen_inv, d.group_gen_inv.l, 32);
    memcpy(size_inv, d.size_inv.l, 32);
    memcpy(generator_inv, d.generator_inv.l, 32);
    return 1;
}

/* domain/utils.rs compute_powers_serial: [1, g, g^2, ...] */
static fr_t *compute_powers_serial(size_t size, const fr_t *root) {
    fr_t *v = (fr_t *)malloc(sizeof(fr_t) * (size ? size : 1));
    fr_t value = fr_R;
    for (size_t i = 0; i < size; i++) {
        v[i] = value;
        fr_mul(&value, &value, root);
    }
    return v;
}
/* radix2/fft.rs:76-138 roots_of_unity: values are [1, g, ..., g^(n/2-1)] in both cfg branches;
 * the parallel branch only changes how they are computed, so threads>1 splits the table in chunks. */
static fr_t *roots_of_unity(const domain_t *d, const fr_t *root, int threads) {
    size_t n = d->size / 2;
    if (threads <= 1 || n < 4096) return compute_powers_serial(n, root);
    fr_t *v = (fr_t *)malloc(sizeof(fr_t) * n);
    size_t chunk = (n + (size_t)threads * 4 - 1) / ((size_t)threads * 4);
    size_t nchunks = (n + chunk - 1) / chunk;
#pragma omp parallel for num_threads(threads)
    for (size_t c = 0; c < nchunks; c++) {
        size_t lo = c * chunk, hi = lo + chunk < n ? lo + chunk : n;
        uint64_t e[1] = {lo};
        fr_t value;
        fr_pow(&value, root, e, 1);
        for (size_t i = lo; i < hi; i++) {
            v[i] = value;
            fr_mul(&value, &value, root);
        }
    }
    return v;
}
/* radix2/fft.rs:249-260 */
static inline uint64_t bitrev(uint64_t a, uint32_t log_len) {
    uint64_t r = 0;
    for (uint32_t i = 0; i < 64; i++) r |= ((a >> i) & 1) << (63 - i);
    return log_len ? r >> (64 - log_len) : 0;
}
static void derange(fr_t *xi, size_t n, uint32_t log_len) {
    if (n < 3) return;
    for (uint64_t idx = 1; idx < (uint64_t)n - 1; idx++) {
        uint64_t ridx = bitrev(idx, log_len);
        if (idx < ridx) {
            fr_t t = xi[idx];
            xi[idx] = xi[ridx];
            xi[ridx] = t;
        }
    }
}
/* radix2/fft.rs:140-203.  threads == 1: the not(parallel) branch with root compaction;
 * threads > 1: the parallel branch (index = nchunks * chunk_index, butterflies split across threads). */
static void io_helper(const domain_t *d, fr_t *xi, size_t n, const fr_t *root, int threads) {
    fr_t *roots = roots_of_unity(d, root, threads);
    size_t root_len = d->size / 2;
    size_t gap = n / 2;
    while (gap > 0) {
        size_t chunk_size = 2 * gap;
        size_t nchunks = n / chunk_size;
        if (threads <= 1) {
            for (size_t c = 0; c < nchunks; c++) {
                fr_t *lo = xi + c * chunk_size, *hi = lo + gap;
                for (size_t idx = 0; idx < gap; idx++) {
                    fr_t neg;
                    fr_sub(&neg, &lo[idx], &hi[idx]);
                    fr_add(&lo[idx], &lo[idx], &hi[idx]);
                    hi[idx] = neg;
                    fr_mul(&hi[idx], &hi[idx], &roots[idx]);
                }
            }
            for (size_t i = 1; i < root_len / 2; i++) roots[i] = roots[i * 2];
            root_len /= 2;
        } else {
#pragma omp parallel for num_threads(threads) schedule(static)
            for (size_t b = 0; b < n / 2; b++) {
                size_t c = b / gap, idx = b % gap;
                fr_t *lo = xi + c * chunk_size, *hi = lo + gap;
                fr_t neg;
                fr_sub(&neg, &lo[idx], &hi[idx]);
                fr_add(&lo[idx], &lo[idx], &hi[idx]);
                hi[idx] = neg;
                fr_mul(&hi[idx], &hi[idx], &roots[nchunks * idx]);
            }
        }
        gap /= 2;
    }
    free(roots);
}
/* radix2/fft.rs:205-234 */
static void oi_helper(const domain_t *d, fr_t *xi, size_t n, const fr_t *root, int threads) {
    fr_t *roots = roots_of_unity(d, root, threads);
    size_t gap = 1;
    while (gap < n) {
        size_t chunk_size = 2 * gap;
        size_t nchunks = n / chunk_size;
#pragma omp parallel for num_threads(threads) schedule(static) if (threads > 1)
        for (size_t b = 0; b < n / 2; b++) {
            size_t c = b / gap, idx = b % gap;
            fr_t *lo = xi + c * chunk_size, *hi = lo + gap;
            fr_mul(&hi[idx], &hi[idx], &roots[nchunks * idx]);
            fr_t neg;
            fr_sub(&neg, &lo[idx], &hi[idx]);
            fr_add(&lo[idx], &lo[idx], &hi[idx]);
            hi[idx] = neg;
        }
        gap *= 2;
    }
    free(roots);
}
/* domain/mod.rs:93-126 distribute_powers_and_mul_by_const: coeff[i] *= c * g^i */
static void distribute_powers_and_mul_by_const(fr_t *coeffs, size_t n, const fr_t *g, const fr_t *c, int threads) {
    if (threads <= 1) {
        fr_t pw = *c;
        for (size_t i = 0; i < n; i++) {
            fr_mul(&coeffs[i], &coeffs[i], &pw);
            fr_mul(&pw, &pw, g);
        }
        return;
    }
    size_t per = n / (size_t)threads;
    if (per < 1024) per = 1024;
    size_t nchunks = (n + per - 1) / per;
#pragma omp parallel for num_threads(threads)
    for (size_t ch = 0; ch < nchunks; ch++) {
        size_t lo = ch * per, hi = lo + per < n ? lo + per : n;
        uint64_t e[1] = {lo};
        fr_t pw;
        fr_pow(&pw, g, e, 1);
        fr_mul(&pw, &pw, c);
        for (size_t i = lo; i < hi; i++) {
            fr_mul(&coeffs[i], &coeffs[i], &pw);
            fr_mul(&pw, &pw, g);
        }
    }
}
/* radix2/fft.rs:22-24,37-51 in_order_fft_in_place: io_helper(group_gen) then derange */
static void in_order_fft(const domain_t *d, fr_t *x, int threads) {
    io_helper(d, x, d->size, &d->group_gen, threads);
    derange(x, d->size, d->log_size);
}
/* radix2/fft.rs:26-29,56-70 */
static void in_order_ifft(const domain_t *d, fr_t *x, int threads) {
    derange(x, d->size, d->log_size);
    oi_helper(d, x, d->size, &d->group_gen_inv, threads);
#pragma omp parallel for num_threads(threads) if (threads > 1)
    for (size_t i = 0; i < d->size; i++) fr_mul(&x[i], &x[i], &d->size_inv);
}
/* radix2/fft.rs:31-35 */
static void in_order_coset_ifft(const domain_t *d, fr_t *x, int threads) {
    derange(x, d->size, d->log_size);
    oi_helper(d, x, d->size, &d->group_gen_inv, threads);
    distribute_powers_and_mul_by_const(x, d->size, &d->generator_inv, &d->size_inv, threads);
}
/* domain/mod.rs:139-142 */
static void coset_fft(const domain_t *d, fr_t *x, int threads) {
    fr_t one = fr_R;
    distribute_powers_and_mul_by_const(x, d->size, &d->generator, &one, threads);
    in_order_fft(d, x, threads);
}
/* The four in-place transforms of EvaluationDomain on a vector that already has domain size
 * (radix2/mod.rs:99-117 resizes with zeros first; callers here pass full-size buffers).
 * data: 2^log_d Montgomery Fr elements, natural order in and out. */
EXPORT int orc_ntt(uint64_t *data, unsigned log_d, int inverse, int coset, int threads) {
    domain_t d;
    if (!domain_new(&d, (size_t)1 << log_d)) return 0;
    fr_t *x = (fr_t *)data;
    if (!inverse && !coset) in_order_fft(&d, x, threads);
    else if (!inverse && coset) coset_fft(&d, x, threads);
    else if (inverse && !coset) in_order_ifft(&d, x, threads);
    else in_order_coset_ifft(&d, x, threads);
    return 1;
}
/* radix2/mod.rs:389-427: the reference test-suite's own serial CLRS radix-2 FFT (bit-reverse, then DIT with
 * per-stage w_m = omega^(n/2m)); a second, independent algorithm to pin the transform. */
EXPORT int orc_serial_radix2_fft(uint64_t *data, unsigned log_n, int inverse) {
    domain_t d;
    size_t n = (size_t)1 << log_n;
    if (!domain_new(&d, n)) return 0;
    fr_t *a = (fr_t *)data;
    fr_t omega = inverse ? d.group_gen_inv : d.group_gen;
    for (uint64_t k = 0; k < n; k++) {
        uint64_t rk = bitrev(k, log_n);
        if (k < rk) {
            fr_t t = a[k];
            a[k] = a[rk];
            a[rk] = t;
        }
    }
    size_t m = 1;
    for (unsigned s = 0; s < log_n; s++) {
        uint64_t e[1] = {n / (2 * m)};
        fr_t w_m;
        fr_pow(&w_m, &omega, e, 1);
        for (size_t k = 0; k < n; k += 2 * m) {
            fr_t w = fr_R;
            for (size_t j = 0; j < m; j++) {
                fr_t t = a[k + j + m];
                fr_mul(&t, &t, &w);
                fr_t tmp;
                fr_sub(&tmp, &a[k + j], &t);
                a[k + j + m] = tmp;
                fr_add(&a[k + j], &a[k + j], &t);
                fr_mul(&w, &w, &w_m);
            }
        }
        m *= 2;
    }
    if (inverse)
        for (size_t i = 0; i < n; i++) fr_mul(&a[i], &a[i], &d.size_inv);
    return 1;
}
/* Horner evaluation of a coefficient vector at x (polynomial evaluate, used by test_fft_correctness) */
EXPORT void orc_poly_eval(uint64_t *out, const uint64_t *coeffs, size_t n, const uint64_t *x) {
    fr_t acc;
    memset(&acc, 0, sizeof acc);
    for (size_t i = n; i-- > 0;) {
        fr_mul(&acc, &acc, (const fr_t *)x);
        fr_add(&acc, &acc, (const fr_t *)(coeffs + 4 * i));
    }
    memcpy(out, acc.l, 32);
}
EXPORT void orc_fr_pow_u64(uint64_t *out, const uint64_t *a, uint64_t e) {
    uint64_t ee[1] = {e};
    fr_pow((fr_t *)out, (const fr_t *)a, ee, 1);
}
/* domain/mod.rs:184-191 + radix2/mod.rs:191-193: evals[i] *= 1 / (g^size - 1) */
EXPORT int orc_divide_by_vanishing_on_coset(uint64_t *data, unsigned log_d, int threads) {
    domain_t d;
    if (!domain_new(&d, (size_t)1 << log_d)) return 0;
    uint64_t e[1] = {d.size};
    fr_t z, zi, one = fr_R;
    fr_pow(&z, &d.generator, e, 1);
    fr_sub(&z, &z, &one);
    fr_inv(&zi, &z);
    fr_t *x = (fr_t *)data;
#pragma omp parallel for num_threads(threads) if (threads > 1)
    for (size_t i = 0; i < d.size; i++) fr_mul(&x[i], &x[i], &zi);
    return 1;
}

#include "czk_oracle_groth16.inc"
#include "czk_oracle_mixed.inc"
#include "czk_oracle_plonk.inc"
