// The per-thread work of one NTT pass on a tile: register-resident radix-8 butterfly phases.
//
// A pass covers r consecutive stages of the radix-2 transform on a tile of 2^r rows x 2^cl columns (2^(r+cl) <= 2048
// elements; row m, column c is element (base | m << L | c) of the vector, L = log_d - s - r, s = the pass's first
// stage).  Inside a pass the stages are grouped into PHASES of up to three: a thread takes the 8 tile elements that
// differ in three consecutive row bits into registers, runs the 12 (or 8, or 4) butterflies of the phase on them and
// puts them back; between two phases the threads exchange elements through the tile in shared memory - one round trip
// per three stages instead of one per stage.
//
// Both in-place radix-2 orders are built from the same index pairs:
//   DIF (natural order in, bit-reversed out):  stages by decreasing span,  (a, b) -> (a + b, (a - b) w)
//   DIT (bit-reversed in, natural out):        stages by increasing span,  (a, b) -> (a + w b, a - w b)
// with the same twiddle for the pair (i, i + h): w = omega^((i mod h) * D / 2h).  An inverse transform uses
// omega^-e = -tw[D/2 - e] (the table holds omega^k for 0 <= k <= D/2), the sign folded into the butterfly, so it
// costs nothing.  Pairing an inverse DIF with a forward DIT (iFFT then coset FFT, mpc-snarks/src/groth/r1cs_to_qap.rs:
// 85-90) needs no bit-reversal pass at all: the scaling D^-1 g^i between them is applied where the DIT loads.
//
// This header compiles for the host too (tests/emu): the CPU test tier runs every phase of every pass thread by thread
// against the CPU restatement of radix2/fft.rs kept with the tests, which checks all the index arithmetic without a GPU.
#pragma once
#include "fp.cuh"

namespace czk {

#ifndef NTT_TILE_LOG_CFG
#define NTT_TILE_LOG_CFG 11
#endif
constexpr int NTT_TILE_LOG = NTT_TILE_LOG_CFG;                // 2048 elements = 64 KB of shared memory per block
constexpr int NTT_TILE_THREADS = (1 << NTT_TILE_LOG) / 8;     // 8 elements per thread
constexpr int NTT_BLOCKS_PER_SM = 512 / NTT_TILE_THREADS;     // 128 registers per thread: 512 threads per SM
constexpr int NTT_MAX_BATCH = 8;       // vectors transformed by one grid

struct NttPass {
    int s;   // first stage of the pass
    int r;   // stages in the pass (>= 3 for every pass of a transform with log_d >= 3)
    int cl;  // log2 of the contiguous columns per tile row
};
struct NttPlan {
    int npass;
    NttPass pass[8];
};
// DIF order (the DIT transform walks the same passes backwards).  The last pass is the contiguous one.
// (tile_log: the tile size; the device always uses NTT_TILE_LOG, the host emulation tests also run smaller tiles so that
// multi-pass plans are exercised at sizes a CPU finishes quickly)
inline NttPlan ntt_make_plan(int log_d, int tile_log = NTT_TILE_LOG) {
    NttPlan p{};
    int last = log_d < tile_log ? log_d : tile_log;
    int rem = log_d - last;
    const int front_cap = tile_log - 3;  // a front pass keeps at least 8 contiguous elements (256 bytes) per tile row
    int nfront = (rem + front_cap - 1) / front_cap;
    int s = 0;
    for (int i = 0; i < nfront; i++) {
        int r = rem / nfront + (i < rem % nfront ? 1 : 0);
        if (r < 3) {  // every pass needs three row bits for its register phases: borrow stages from the last pass
            last -= 3 - r;
            r = 3;
        }
        int L = log_d - s - r;
        int cl = tile_log - r;
        if (cl > L) cl = L;
        p.pass[p.npass++] = NttPass{s, r, cl};
        s += r;
    }
    p.pass[p.npass++] = NttPass{s, last, 0};
    return p;
}

// What multiplies an element when a pass loads it (pre) or stores it (post): nothing, a constant, or
// lo[i & (2^lo_log - 1)] * hi[i >> lo_log] for the element's NATURAL index i (= the bit reversal of its position when
// the data is in bit-reversed order at that point).
struct NttScale {
    int mode = 0;  // 0 none, 1 constant, 2 two-level power table, 3 one factor per POSITION read from `lo` (already permuted)
    int bitrev = 0;
    uint32_t c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint32_t* lo = nullptr;
    const uint32_t* hi = nullptr;
    int lo_log = 0;
};

struct NttTileGeom {
    int n, s, r, cl, L;
    size_t base;  // global index of the tile's element 0
};

CZK_HD int ntt_num_phases(int r) { return (r + 2) / 3; }
// phase `ph` of a pass: the thread's three element bits are tile bits [kp, kp + 3); the phase's stages are the
// element bits kp + jlo .. kp + jlo + ns - 1
CZK_HD void ntt_phase_geom(int r, int cl, bool dit, int ph, int& kp, int& jlo, int& ns) {
    const int nph = ntt_num_phases(r), last_ns = r - 3 * (nph - 1);
    if (!dit) {  // from the top row bit down
        if (ph < nph - 1) {
            ns = 3;
            kp = cl + r - 3 * (ph + 1);
            jlo = 0;
        } else {
            ns = last_ns;
            kp = cl;
            jlo = 0;
        }
    } else {  // from the bottom row bit up
        if (ph < nph - 1) {
            ns = 3;
            kp = cl + 3 * ph;
            jlo = 0;
        } else {
            ns = last_ns;
            kp = cl + r - 3;
            jlo = 3 - last_ns;
        }
    }
}

CZK_HD Fr ntt_ld_words(const uint32_t* p) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = p[i];
    return r;
}
CZK_HD uint32_t ntt_bitrev(uint32_t v, int bits) {
    uint32_t r = 0;
#ifdef __CUDA_ARCH__
    r = bits ? (__brev(v) >> (32 - bits)) : 0u;
#else
    for (int i = 0; i < bits; i++) r |= ((v >> i) & 1u) << (bits - 1 - i);
#endif
    return r;
}

template <class TW>
CZK_HD Fr ntt_scale_factor(const NttScale& sc, size_t pos, int n, TW ldtw) {
    if (sc.mode == 1) return ntt_ld_words(sc.c);
    if (sc.mode == 3) return ldtw(sc.lo, pos);
    const size_t i = sc.bitrev ? (size_t)ntt_bitrev((uint32_t)pos, n) : pos;
    return Fr::mul(ldtw(sc.lo, i & (((size_t)1 << sc.lo_log) - 1)), ldtw(sc.hi, i >> sc.lo_log));
}

// one stage on element bit J of the 8 registers: 4 butterflies
template <int J, bool DIT, class TW>
CZK_HD void ntt_stage(Fr* x, bool inverse, const uint32_t* tw, size_t lowpart, int n, int lq, TW ldtw) {
    const size_t half = (size_t)1 << (n - 1);
    const int sh = n - 1 - lq - J;  // the pair (i, i + h), h = 2^(lq + J), uses omega^((i mod h) << sh)
    // Unit twiddles need no product.  They are exactly the pairs whose lower index is 0 mod h: the whole stage when
    // h = 1 (lq + J == 0), and the butterflies with no lower element bits set when lq == 0.
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (k & (1 << J)) continue;
        const int klow = k & ((1 << J) - 1);
        const size_t ex = (lowpart << sh) + ((size_t)klow << (n - 1 - J));
        const bool unit = lq == 0 && klow == 0;  // lowpart = 0 when lq == 0
        Fr& a = x[k];
        Fr& b = x[k | (1 << J)];
        if (!DIT) {
            const Fr sum = Fr::add(a, b);
            Fr d = inverse ? Fr::sub(b, a) : Fr::sub(a, b);  // inverse: (a - b) * -tw[half - e] = (b - a) * tw[half - e]
            if (unit) d = inverse ? Fr::sub(a, b) : d;       // tw[half] = -1
            else d = Fr::mul(d, ldtw(tw, inverse ? half - ex : ex));
            a = sum;
            b = d;
        } else {
            Fr t = b;
            if (!unit) t = Fr::mul(b, ldtw(tw, inverse ? half - ex : ex));
            const bool flip = inverse && !unit;  // inverse: w b = -(tw[half - e] b)
            const Fr p = Fr::add(a, t), m = Fr::sub(a, t);
            a = flip ? m : p;
            b = flip ? p : m;
        }
    }
}

// One thread's share of one phase.  Tile: element accessor with load(e) / store(e, v); ldtw(table, index) loads one
// Montgomery Fr from a device table.  first / last: whether this is the first phase of the first pass / the last
// phase of the last pass of the transform (where pre / post apply).
// DIT and SCALE are compile-time so that a kernel carries one butterfly form and, unless it is the scaling variant, no
// scaling code (the inlined products are the bulk of the instruction footprint).
template <bool DIT, bool SCALE, class Tile, class TW>
CZK_HD void ntt_phase_thread(Tile& tile, unsigned u, const NttTileGeom& g, bool inverse, int kp, int jlo, int ns,
                             const uint32_t* tw, bool first, bool last, const NttScale& pre, const NttScale& post, TW ldtw) {
    const unsigned e0 = ((u >> kp) << (kp + 3)) | (u & ((1u << kp) - 1u));
    Fr x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = tile.load(e0 | ((unsigned)k << kp));
    const size_t g0 = g.base | ((size_t)(e0 >> g.cl) << g.L) | (size_t)(e0 & ((1u << g.cl) - 1u));
    const int lq = g.L + kp - g.cl;  // log2 of the distance, in the vector, between the thread's consecutive elements
    if (SCALE && first && pre.mode) {
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = Fr::mul(x[k], ntt_scale_factor(pre, g0 + ((size_t)k << lq), g.n, ldtw));
    }
    const size_t lowpart = g0 & (((size_t)1 << lq) - 1);
    if (!DIT) {
        if (jlo <= 2 && 2 < jlo + ns) ntt_stage<2, DIT>(x, inverse, tw, lowpart, g.n, lq, ldtw);
        if (jlo <= 1 && 1 < jlo + ns) ntt_stage<1, DIT>(x, inverse, tw, lowpart, g.n, lq, ldtw);
        if (jlo <= 0 && 0 < jlo + ns) ntt_stage<0, DIT>(x, inverse, tw, lowpart, g.n, lq, ldtw);
    } else {
        if (jlo <= 0 && 0 < jlo + ns) ntt_stage<0, DIT>(x, inverse, tw, lowpart, g.n, lq, ldtw);
        if (jlo <= 1 && 1 < jlo + ns) ntt_stage<1, DIT>(x, inverse, tw, lowpart, g.n, lq, ldtw);
        if (jlo <= 2 && 2 < jlo + ns) ntt_stage<2, DIT>(x, inverse, tw, lowpart, g.n, lq, ldtw);
    }
    if (SCALE && last && post.mode) {
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = Fr::mul(x[k], ntt_scale_factor(post, g0 + ((size_t)k << lq), g.n, ldtw));
    }
#pragma unroll
    for (int k = 0; k < 8; k++) tile.store(e0 | ((unsigned)k << kp), x[k]);
}

// The combining step of a mixed-radix transform over 3 M points (ntt.cu, k_mr_combine): from the three radix-2 outputs at
// index j (y[r] = Y_r[j]), w = omega^j and zeta = omega^M, X[j + s M] = y0 + zeta^s w y1 + zeta^(2 s) w^2 y2 for s = 0, 1, 2.
// scale: multiply everything by c (3^-1 on the inverse).  Two products by zeta: zeta^2 t = -(t + zeta t).
CZK_HD void ntt_mixed_combine3(const Fr y[3], const Fr& w, const Fr& zeta, const Fr& c, bool scale, Fr out[3]) {
    Fr y0 = y[0], y1 = y[1], y2 = y[2];
    if (scale) {
        y0 = Fr::mul(y0, c);
        y1 = Fr::mul(y1, c);
        y2 = Fr::mul(y2, c);
    }
    const Fr t1 = Fr::mul(y1, w), t2 = Fr::mul(y2, Fr::mul(w, w));
    const Fr z1 = Fr::mul(t1, zeta), z2 = Fr::mul(t2, zeta);
    const Fr zz1 = Fr::neg(Fr::add(t1, z1)), zz2 = Fr::neg(Fr::add(t2, z2));
    out[0] = Fr::add(y0, Fr::add(t1, t2));
    out[1] = Fr::add(y0, Fr::add(z1, zz2));
    out[2] = Fr::add(y0, Fr::add(zz1, z2));
}

}  // namespace czk
