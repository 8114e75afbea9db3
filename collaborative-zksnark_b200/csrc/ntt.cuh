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
//   * one twiddle table w^k, 0 <= k <= D/2, per domain size, built once on the device and kept in HBM (the reference
//     recomputes it every call, fft.rs:144,206); inverse twiddles are read from the same table as w^-k = -w^(D/2-k)
//     with the sign folded into the butterfly;
//   * log D stages grouped into passes; a pass moves a 2^r x 2^cl tile (r stages deep, 2^cl consecutive columns wide,
//     2048 elements = 64 KB) into shared memory with bulk asynchronous copies (TMA: one cp.async.bulk per tile row
//     completing on an mbarrier) and back with bulk stores; inside the pass every thread keeps 8 elements in registers
//     for three stages at a time (radix-8 phases, ntt_tile.cuh), so an element crosses shared memory once per three
//     stages; `count` vectors share one grid (blockIdx.y);
//   * both in-place orders exist: DIF (natural -> bit-reversed) and DIT (bit-reversed -> natural).  A standalone
//     transform is DIF + one bit-reversal swap that also applies the inverse / coset scaling; the witness map's
//     iFFT -> coset FFT pair is inverse DIF -> forward DIT with the D^-1 g^i scaling fused into the DIT's first loads
//     and no reordering pass at all.
// Measured (B200, 2^21, profiles/): 0.50 ms per transform in a batch of 6 (the butterfly is 367 SASS instructions, every
// integer instruction holds a dispatch port for ~1.9 cycles: floor 0.42 ms), 64 B/element/pass of HBM traffic.
#pragma once
#include <cuda_runtime.h>
#include "fp.cuh"
#include "ntt_tile.cuh"

namespace czk {

// table[k] = c * base^k, k < n (Montgomery)
cudaError_t ntt_build_powers(uint32_t* table, const uint64_t base[4], const uint64_t c[4], size_t n, cudaStream_t st);
// All the passes of one in-place transform over `count` <= NTT_MAX_BATCH vectors in one grid per pass.  tw: omega^k for
// 0 <= k <= D/2.  dit = false: natural order in, bit-reversed out; dit = true: bit-reversed in, natural out.  pre / post:
// scaling fused into the first pass's loads / the last pass's stores.
cudaError_t ntt_run_tiles(uint32_t* const* data, int count, const uint32_t* tw, int log_d, bool inverse, bool dit, const NttScale& pre,
                          const NttScale& post, cudaStream_t st);
// data[i] *= lo[i & (2^lo_log - 1)] * hi[i >> lo_log]
cudaError_t ntt_scale_by_powers(uint32_t* data, const uint32_t* lo, const uint32_t* hi, int lo_log, int log_d, cudaStream_t st);
// in-place bit reversal; mode 0: none, 1: times constant c, 2: times lo/hi power tables (hi carries the constant)
cudaError_t ntt_bitrev_scale(uint32_t* data, int log_d, int mode, const uint64_t c[4], const uint32_t* lo,
                             const uint32_t* hi, int lo_log, cudaStream_t st);

// mixed radix (N = 3 M, M = 2^k): out[r M + a] = in[3 a + r]; and the combining pass after the three radix-2 transforms
// (wpow[j] = w^j, j < M; zeta = w^M; c: optional constant multiplied into every output)
cudaError_t ntt_mixed_split(const uint32_t* in, uint32_t* out, size_t M, cudaStream_t st);
cudaError_t ntt_mixed_combine(const uint32_t* y, uint32_t* out, const uint32_t* wpow, const uint64_t zeta[4], const uint64_t c[4], size_t M,
                              cudaStream_t st);

}  // namespace czk
