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
    unsigned nwin;     // windows
    unsigned nb;       // buckets per window = 2^(c-1)
    unsigned chunk;    // buckets per thread in the reduce kernel
};

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
    return cfg;
}

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
    uint32_t* queue = nullptr;     // {item count, queue head}
    int sm_count = 148;
    int seg_point_words = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // accumulate start/stop, whole MSM start/stop
    size_t cap_n = 0;
    size_t cap_buckets = 0;
    int point_words = 0;
};

// curve: 1 = G1 (Fq, 12 words per coordinate), 2 = G2 (Fq2, 24 words per coordinate)
// bases: n affine points, x | y, Montgomery, 16-byte aligned; inf: n bytes or nullptr.
// scalars: n x 8 words.  winsum_out: device buffer of cfg.nwin XYZZ points.
cudaError_t msm_run(int curve, const uint32_t* bases, const uint8_t* inf, const uint32_t* scalars, bool scalars_mont,
                    size_t n, const MsmConfig& cfg, MsmWorkspace& ws, cudaStream_t st);

size_t msm_point_words(int curve);  // 4 coordinates

// test / synthetic-input helper: out[i] = (k0 + i * kstep) * base as affine points (x | y)
// (step_xy = kstep * base, affine, device memory)
cudaError_t ec_gen_progression_dev(int curve, uint32_t* out_xy, const uint32_t* base_xy, const uint32_t* step_xy,
                                   const uint64_t k0_canon[4], const uint64_t kstep_canon[4], size_t n, cudaStream_t st);

}  // namespace czk
