// Radix-2 NTT / iNTT over BLS12-377 Fr on the device.
//
// Replaces Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place
// (algebra/poly/src/domain/radix2/mod.rs:99-117, radix2/fft.rs:22-260,
//  domain/mod.rs:93-142).  Semantics kept bit for bit: natural order in and out,
//  out[i] = sum_j in[j] w^(ij) with w = get_root_of_unity(D) (large-subgroup branch:
//  11^((r-1)/D), NOT the two-adic root), coset shift g = 22, inverse scaled by D^-1,
//  coset inverse scaled by D^-1 g^-i.
//
// Device algorithm (the reference's serial DIF + derange is not copied):
//   * one twiddle table w^k, k < D/2, per domain size, built once on the device and kept in HBM
//     (the reference recomputes it every call, fft.rs:144,206); inverse twiddles are read from the
//     same table as w^-k = -w^(D/2-k);
//   * decimation-in-frequency in log D stages grouped into passes; a pass keeps a 2^r x 2^cl tile
//     (r butterfly stages deep, 2^cl consecutive columns wide so every global row segment is a
//     contiguous 2^cl * 32 B run) in shared memory, limb-major so that consecutive threads hit
//     consecutive banks;
//   * a final in-place bit-reversal swap that also applies the inverse / coset scaling.
#pragma once
#include <cuda_runtime.h>
#include "fp.cuh"

namespace czk {

constexpr int NTT_TILE_LOG = 11;  // 2048 elements = 64 KB of shared memory per block
constexpr int NTT_THREADS = 256;

struct NttPlanPass {
    int s;   // first stage of the pass
    int r;   // stages in the pass
    int cl;  // log2 of the contiguous columns per tile row
};

struct NttPlan {
    int npass;
    NttPlanPass pass[8];
};

inline NttPlan ntt_make_plan(int log_d) {
    NttPlan p{};
    int last = log_d < NTT_TILE_LOG ? log_d : NTT_TILE_LOG;
    int rem = log_d - last;
    int nfront = (rem + 7) / 8;
    int s = 0;
    for (int i = 0; i < nfront; i++) {
        int r = rem / nfront + (i < rem % nfront ? 1 : 0);
        int L = log_d - s - r;
        int cl = NTT_TILE_LOG - r;
        if (cl > L) cl = L;
        p.pass[p.npass++] = NttPlanPass{s, r, cl};
        s += r;
    }
    p.pass[p.npass++] = NttPlanPass{s, last, 0};
    return p;
}

// table[k] = c * base^k, k < n (Montgomery)
cudaError_t ntt_build_powers(uint32_t* table, const uint64_t base[4], const uint64_t c[4], size_t n, cudaStream_t st);
cudaError_t ntt_run_passes(uint32_t* data, const uint32_t* tw, int log_d, bool inverse, cudaStream_t st);
// data[i] *= lo[i & (2^lo_log - 1)] * hi[i >> lo_log]
cudaError_t ntt_scale_by_powers(uint32_t* data, const uint32_t* lo, const uint32_t* hi, int lo_log, int log_d, cudaStream_t st);
// in-place bit reversal; mode 0: none, 1: times constant c, 2: times lo/hi power tables (hi carries the constant)
cudaError_t ntt_bitrev_scale(uint32_t* data, int log_d, int mode, const uint64_t c[4], const uint32_t* lo,
                             const uint32_t* hi, int lo_log, cudaStream_t st);

}  // namespace czk
