// Variable-base multi-scalar multiplication on BLS12-377 G1 / G2.
//
// Replaces VariableBaseMSM::multi_scalar_mul (algebra/ec/src/msm/variable_base.rs:12-106) and its
// wrapper AffineCurve::multi_scalar_mul (algebra/ec/src/lib.rs:302-311: Fr -> BigInt per scalar).
// Same function - sum_i s_i P_i over min(len) terms, infinity bases and zero scalars contribute
// nothing - computed by a device bucket method that shares nothing with the reference's loop
// structure:
//   1. prepare   : scalar from Montgomery form, balanced signed c-bit digits, per-bucket histogram
//   2. scan      : exclusive prefix sum of the histogram -> bucket offsets
//   3. scatter   : counting sort of (point index, sign) by (window, |digit|)
//   4. accumulate: one thread per bucket, XYZZ mixed additions over its sorted run
//   5. reduce    : per window sum_b b * B_b  by chunked running sums, then a block tree sum
//   6. host tail : sum_w 2^(c w) W_w (c doublings per window, variable_base.rs:92-105) and one
//                  affine normalisation, on the CPU - it is O(windows), strictly serial.
// The window size c is chosen for the GPU, not by the reference's ln_without_floats heuristic;
// the group element returned is the same, and it is returned in affine-normalised form.
#pragma once
#include <cuda_runtime.h>
#include "ec.cuh"
#include "msm_digits.cuh"

namespace czk {

struct MsmConfig {
    unsigned c;        // window bits
    unsigned nwin;     // digit windows W = ceil(254 / c)
    unsigned nb;       // buckets per bucket-window = 2^(c-1)
    unsigned chunk;    // buckets per thread in the reduce kernel
    // Merged windows: with the multiples 2^(c w) P_i precomputed (czk_bases_precompute) every digit of every window
    // feeds ONE set of 2^(c-1) buckets - the (point, window) pair just selects a different table entry - so the
    // bucket reduction runs once instead of W times, larger c (fewer additions) becomes affordable, and the host
    // tail needs no doublings.  bwin = number of bucket sets (1 when merged, W otherwise).
    unsigned merged = 0;
    unsigned bwin = 0;
    size_t table_stride = 0;  // merged: points per window slab of the table (= length of the uploaded base array)
    size_t table_off = 0;     // merged: index of the first base of this MSM inside a slab
    // words from one point record of `bases` to the next (0: packed x | y).  G1 table records are padded from 96 to 128
    // bytes so that a gather touches exactly one 128-byte line: with packed records half of them straddle two lines and
    // the measured DRAM traffic of round 0 was 1.5 lines per gather (profiles/r2_summary.md)
    unsigned base_stride = 0;
};

// window size for the merged form: accumulation is n * ceil(254/c) additions, reduction ~3.3 * 2^(c-1)
inline unsigned msm_merged_window(size_t n) {
    unsigned lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    // only sizes whose top window is (nearly) full: its digits land in the low buckets, and a window holding b bits
    // puts n / 2^b extra points into each of them (c = 18, 19, 21 leave 1, 6 and 1 bits)
    unsigned c = lg >= 23 ? 20 : (lg >= 18 ? 17 : (lg >= 14 ? 16 : (lg > 10 ? lg - 2 : 8)));
    return c;
}
inline MsmConfig msm_merged_config(unsigned c, size_t stride, size_t off, unsigned base_stride = 0) {
    MsmConfig cfg;
    cfg.c = c;
    cfg.nwin = msm_num_windows(c);
    cfg.nb = 1u << (c - 1);
    // one bucket set only: the chunked running sums are latency-bound (few threads, ~30 us per addition with one warp
    // per scheduler), so use small chunks here - ~16K threads - unlike the W-window form where chunk 4 doubles the work
    cfg.chunk = 4;
    while (cfg.nb / cfg.chunk > 16384 && cfg.chunk < 32) cfg.chunk <<= 1;
    if (cfg.chunk > cfg.nb) cfg.chunk = cfg.nb;
    cfg.merged = 1;
    cfg.bwin = 1;
    cfg.table_stride = stride;
    cfg.table_off = off;
    cfg.base_stride = base_stride;
    return cfg;
}

inline MsmConfig msm_choose_config(size_t n) {
    // c ~ log2(n) - 4, clamped; measured trade-off between N*W mixed adds and 2^c*W bucket work
    unsigned lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    // 2^(c-1) buckets per window: c = lg - 4 balances N*W mixed additions against 2^c*W bucket-reduction additions
    // up to 2^19; from 2^20 on one bit fewer halves the reduction for +6% accumulation (measured, profiles/)
    unsigned c = lg > 6 ? lg - 4 : 2;
    if (lg == 20) c = 15;
    if (c < 2) c = 2;
    if (c > 16) c = 16;
    // avoid window sizes whose top window holds only a few bits of the 253-bit scalar: its few buckets
    // would each receive n / 2^bits points.  (c = 16, 15, 11 and the small sizes are balanced.)
    if (c == 14) c = 15;
    if (c == 12 || c == 13) c = 11;
    if (c == 10 || c == 9) c = 8;
    MsmConfig cfg;
    cfg.c = c;
    cfg.nwin = msm_num_windows(c);
    cfg.nb = 1u << (c - 1);
    // buckets per thread in the chunked running sums.  Each chunk pays a ~1.5 log2(nb)-addition scalar multiple for
    // its offset, so small chunks double the work (measured: chunk 4 is 3.7 ms slower per 2^20 MSM than 16)
    cfg.chunk = cfg.nb >= 16 ? 16 : cfg.nb;
    cfg.merged = 0;
    cfg.bwin = cfg.nwin;
    return cfg;
}

// Merged form with enough buckets: the bucket reduction is c bit-slice sums (k_msm_bit_sums) and winsum holds c points
// S_0 .. S_{c-1} with sum_b (b+1) B_b = sum_j 2^j S_j; otherwise winsum holds one point per bucket window.
constexpr unsigned MSM_BITSUM_THREADS = 2048;  // threads per slice, at least; nb / 32 for larger bucket sets (16 adds each)
inline bool msm_uses_bit_sums(const MsmConfig& cfg) { return cfg.merged && cfg.nb >= 2 * MSM_BITSUM_THREADS; }
inline unsigned msm_bitsum_threads(const MsmConfig& cfg) { return cfg.nb / 32 > MSM_BITSUM_THREADS ? cfg.nb / 32 : MSM_BITSUM_THREADS; }
inline unsigned msm_winsum_points(const MsmConfig& cfg) { return msm_uses_bit_sums(cfg) ? cfg.c : cfg.bwin; }

struct MsmWorkspace {
    uint32_t* scalars = nullptr;   // n x 8 canonical
    uint32_t* hist = nullptr;      // nwin*nb
    uint32_t* offsets = nullptr;   // nwin*nb (exclusive scan, then advanced to bucket ends by the scatter)
    uint32_t* sorted = nullptr;    // n*nwin entries: point index | sign << 31
    uint32_t* buckets = nullptr;   // nwin*nb points, XYZZ
    uint32_t* partial = nullptr;   // nwin*(nb/chunk) points
    uint32_t* winsum = nullptr;    // nwin points
    uint32_t* segcnt = nullptr;    // nwin*nb: segments per bucket
    uint32_t* segoff = nullptr;    // nwin*nb: exclusive scan of segcnt
    uint32_t* segsum = nullptr;    // cap_items points: per-segment sums of multi-segment buckets
    size_t cap_items = 0;
    uint32_t* items = nullptr;     // cap_items x uint4 item descriptors
    uint32_t* queue = nullptr;     // {item count, queue head, heavy count}
    uint32_t* heavy = nullptr;     // cap_items item ids served first
    // batched-affine accumulation (msm_batched.cu): ping-pong point arrays, prefix-product scratch
    uint32_t *bat_a = nullptr, *bat_b = nullptr, *bat_prefix = nullptr;
    size_t cap_bat_a = 0, cap_bat_b = 0, cap_bat_prefix = 0;  // bytes
    bool batched = true;                                       // CZK_BATCHED=0 keeps the XYZZ walk
    bool batched_forced = false;                               // set by czk_msm_set_batched
    bool batched_always = false;                               // czk_msm_set_batched(ctx, 2): ignore the size thresholds (tests)
    int sm_count = 148;
    int seg_point_words = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // accumulate start/stop, whole MSM start/stop
    size_t cap_n = 0;
    size_t cap_sorted_bytes = 0;
    size_t cap_buckets = 0;
    int point_words = 0;
    // the last run's plan (sorted entries, bucket runs, item queue) can serve another base set: see msm_run's reuse_plan
    uint64_t alloc_epoch = 0;                        // bumped whenever a plan buffer is reallocated
};

// curve: 1 = G1 (Fq, 12 words per coordinate), 2 = G2 (Fq2, 24 words per coordinate)
// bases: n affine points, x | y, Montgomery, 16-byte aligned; inf: n bytes or nullptr.
// scalars: n x 8 words.  winsum_out: device buffer of cfg.nwin XYZZ points.
// reuse_plan: the workspace still holds the plan of the previous run, which had the same scalars, n, cfg and infinity flags
// (the caller's claim): skip the digit decomposition and the bucket sort, run accumulation and reduction over `bases`.
cudaError_t msm_run(int curve, const uint32_t* bases, const uint8_t* inf, const uint32_t* scalars, bool scalars_mont,
                    size_t n, const MsmConfig& cfg, MsmWorkspace& ws, cudaStream_t st, bool reuse_plan = false);
// *differ (device word, must be zero on entry) becomes non-zero iff a[i] != b[i] for some i < n
cudaError_t msm_flags_differ(const uint8_t* a, const uint8_t* b, size_t n, uint32_t* differ, cudaStream_t st);

size_t msm_point_words(int curve);  // 4 coordinates

// Batched-affine bucket accumulation (msm_batched.cu).  ends / hist: the counting sort's bucket ends and lengths;
// entries: an upper bound on the sorted entries (n * windows), which sizes the grids; maxlen_dev: DEVICE word holding the
// longest bucket, from which the kernels themselves decide how many halving rounds run.  Writes `buckets` (XYZZ)
// unless an addition without an affine formula was met, in which case *flag becomes non-zero and the caller's XYZZ
// kernel must run.  pa / pb / prefix: scratch of the sizes msm_batched_bytes reports.
size_t msm_batched_bytes(int curve, size_t entries, size_t nb, size_t* pa, size_t* pb, size_t* pre);
// bstride: words between consecutive point records of `bases` (2 W when packed)
cudaError_t msm_batched_accumulate(int curve, const uint32_t* bases, unsigned bstride, const uint32_t* sorted, const uint32_t* ends,
                                   const uint32_t* hist, size_t nb, size_t entries, const uint32_t* maxlen_dev, uint32_t* pa,
                                   uint32_t* pb, uint32_t* prefix, uint32_t* buckets, uint32_t* flag, int sm_count, cudaStream_t st);
// out[i] = 1 / in[i] in Fq (Montgomery), 0 -> 0: the block inversion of the batched path, exposed for its parity test
cudaError_t fq_inverse_batch(const uint32_t* in, uint32_t* out, size_t n, cudaStream_t st);

// table[w * n + i] = 2^(c w) * bases[i] as affine points, w < nwin (slab 0 is a copy of the input)
// tstride: words between consecutive records of the table (>= 2 W)
cudaError_t msm_precompute_table(int curve, uint32_t* table, unsigned tstride, const uint32_t* bases, size_t n, unsigned c, unsigned nwin,
                                 cudaStream_t st);
// record stride (words) of a merged-window table: G1 records are padded to one 128-byte line
inline unsigned msm_table_stride_words(int curve) { return curve == 1 ? 32u : 48u; }

// synthetic-input helper: out[i] = (k0 + i kstep + i^2 kquad) * base as affine points (x | y); quad2_xy = the affine
// point 2 kquad * base (device memory).  The quadratic term matters: with a plain arithmetic progression every
// difference P_i - P_j = (i - j) kstep base repeats, so partial bucket sums coincide far more often than between the
// independent-looking points of a real CRS.
cudaError_t ec_gen_progression_dev(int curve, uint32_t* out_xy, const uint32_t* base_xy, const uint32_t* quad2_xy,
                                   const uint64_t k0_canon[4], const uint64_t kstep_canon[4], const uint64_t kquad_canon[4],
                                   size_t n, cudaStream_t st);

}  // namespace czk
