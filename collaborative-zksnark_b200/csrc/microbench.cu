// Integer-pipe microbenchmarks: the roofline denominators for the field kernels.  MEASURED_PEAKS.json
// carries only HBM and bf16 peaks; these measure IMAD / IMAD.WIDE issue rates and the achieved field
// multiplication and point addition rates at a chosen occupancy.
#include <cuda_runtime.h>
#include "ec.cuh"
#include "launch_count.hpp"

namespace czk {

// kind 0: independent chains of 32x32+64 -> 64 multiply-adds (IMAD.WIDE.U32)
__global__ void k_mb_wide(uint64_t* out, int iters, uint32_t seed) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    uint64_t c0 = a, c1 = b, c2 = a ^ b, c3 = a + b, c4 = 5, c5 = 6, c6 = 7, c7 = 8;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c0) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c1) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c2) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c3) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c4) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c5) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c6) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c7) : "r"(a), "r"(b));
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}
// kind 1: the same work as (mad.lo, mad.hi) pairs = 2 IMAD issue slots per product
__global__ void k_mb_pair(uint64_t* out, int iters, uint32_t seed) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    uint32_t l0 = a, h0 = b, l1 = 1, h1 = 2, l2 = 3, h2 = 4, l3 = 5, h3 = 6;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(l0) : "r"(a), "r"(b));
            asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(h0) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(l1) : "r"(a), "r"(b));
            asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(h1) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(l2) : "r"(a), "r"(b));
            asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(h2) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(l3) : "r"(a), "r"(b));
            asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(h3) : "r"(a), "r"(b));
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((uint64_t)(l0 ^ l1 ^ l2 ^ l3) << 32) | (h0 ^ h1 ^ h2 ^ h3);
}
// kind 5: carry-chained wide multiply-adds (IMAD.WIDE.U32.X), the form the Montgomery rows issue
__global__ void k_mb_wide_carry(uint64_t* out, int iters, uint32_t seed) {
    uint32_t a[12], acc0[12], acc1[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        a[i] = seed * (i + 1) + threadIdx.x;
        acc0[i] = i;
        acc1[i] = blockIdx.x + i;
    }
    uint32_t b0 = seed ^ threadIdx.x, b1 = b0 * 3 + 1, b2 = b0 * 5 + 7, b3 = b0 * 7 + 11;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        chain_mad<12>(acc0, a, b0);
        chain_mad<12>(acc1, a, b1);
        chain_mad<12>(acc0, a, b2);
        chain_mad<12>(acc1, a, b3);
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) r ^= ((uint64_t)acc0[i] << 32) | acc1[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__device__ uint32_t g_mb_in[64];
__device__ uint32_t g_mb_lane_zero[32];
// kind 13: carry-chained wide multiply-adds, accumulators updated in place (operands loaded from memory so that the
// register allocation matches the field kernels': the loop is IMAD.WIDE.U32[.X] + one IADD3.X per chain)
__global__ void k_mb_wide_carry_clean(uint64_t* out, int iters, uint32_t seed) {
    uint32_t a[12], acc[12], acc2[12], b[4];
    int z = g_mb_lane_zero[threadIdx.x & 31];
#pragma unroll
    for (int i = 0; i < 12; i++) a[i] = g_mb_in[i + z], acc[i] = g_mb_in[12 + i + z], acc2[i] = g_mb_in[24 + i + z];
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = g_mb_in[40 + i + z] + seed;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r += 2) {
            chain_mad<12>(acc, a, b[r]);
            acc[11] = addc(acc[11], 0);
            chain_mad<12>(acc2, a, b[r + 1]);
            acc2[11] = addc(acc2[11], 0);
        }
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) r ^= ((uint64_t)acc[i] << 32) | acc2[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// kind 14: wide multiply-adds (multiply pipe) interleaved 1:1 with carry-chained adds (ALU pipe): do the pipes overlap?
__global__ void k_mb_wide_plus_add(uint64_t* out, int iters, uint32_t seed) {
    int z = g_mb_lane_zero[threadIdx.x & 31];
    uint32_t a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = g_mb_in[i + z] + threadIdx.x;
    uint32_t b = g_mb_in[z + 17] + seed;
    uint64_t c0 = a[0], c1 = b, c2 = a[1] ^ b, c3 = a[2] + b;
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = g_mb_in[20 + i + z];
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c0) : "r"(a[4 * k]), "r"(b));
            x[0] = add_cc(x[0], x[4]);
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c1) : "r"(a[4 * k + 1]), "r"(b));
            x[1] = addc_cc(x[1], x[5]);
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c2) : "r"(a[4 * k + 2]), "r"(b));
            x[2] = addc_cc(x[2], x[6]);
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c3) : "r"(a[4 * k + 3]), "r"(b));
            x[3] = addc(x[3], x[7]);
        }
        b += (uint32_t)c3;
    }
    uint64_t r = c0 ^ c1 ^ c2 ^ c3;
#pragma unroll
    for (int i = 0; i < 4; i++) r ^= x[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// kind 6: carry-chained 32-bit adds (IADD3.X)
__global__ void k_mb_addc(uint64_t* out, int iters, uint32_t seed) {
    uint32_t x[12], y[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        x[i] = seed * (i + 1) + threadIdx.x;
        y[i] = blockIdx.x + i * 77;
    }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            x[0] = add_cc(x[0], y[0]);
#pragma unroll
            for (int j = 1; j < 11; j++) x[j] = addc_cc(x[j], y[j]);
            x[11] = addc(x[11], y[11]);
            y[0] = add_cc(y[0], x[0]);
#pragma unroll
            for (int j = 1; j < 11; j++) y[j] = addc_cc(y[j], x[j]);
            y[11] = addc(y[11], x[11]);
        }
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) r ^= ((uint64_t)x[i] << 32) | y[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// kind 7: independent 32-bit IMAD (lo) chains
__global__ void k_mb_imad(uint64_t* out, int iters, uint32_t seed) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    uint32_t c0 = 1, c1 = 2, c2 = 3, c3 = 4, c4 = 5, c5 = 6, c6 = 7, c7 = 8;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c0) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c1) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c2) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c3) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c4) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c5) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c6) : "r"(a), "r"(b));
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c7) : "r"(a), "r"(b));
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}

template <class F>
__global__ void k_mb_mul(uint64_t* out, int iters, uint32_t seed) {
    F x = F::one(), y = F::r2();
    x.l[0] ^= seed + threadIdx.x;
    y.l[1] ^= blockIdx.x;
    for (int i = 0; i < iters; i++) {
        x = F::mul(x, y);
        y = F::mul(y, x);
    }
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < F::N; i++) acc ^= (uint64_t)(x.l[i] ^ y.l[i]) << (i & 31);
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// kind 15: TWO independent Fq product chains per thread - does a second product in flight (4 carry chains instead of 2)
// raise the multiply-pipe utilisation of a lone warp?  (kind 3 reaches 67 % of the 4-warp rate with one warp per scheduler.)
__global__ void __launch_bounds__(256) k_mb_mul_x2(uint64_t* out, int iters, uint32_t seed) {
    Fq x = Fq::one(), y = Fq::r2(), u = Fq::r2(), v = Fq::one();
    x.l[0] ^= seed + threadIdx.x;
    y.l[1] ^= blockIdx.x;
    u.l[2] ^= seed ^ threadIdx.x;
    v.l[3] ^= blockIdx.x + 7u;
    for (int i = 0; i < iters; i++) {
        Fq::mul2(x, y, u, v, x, u);
        Fq::mul2(y, x, v, u, y, v);
    }
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < Fq::N; i++) acc ^= (uint64_t)(x.l[i] ^ y.l[i] ^ u.l[i] ^ v.l[i]) << (i & 31);
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_mb_madd(uint64_t* out, int iters, uint32_t seed) {
    XYZZ<Fq> acc = XYZZ<Fq>::from_affine(Fq::r2(), Fq::one());
    Fq px = Fq::one(), py = Fq::r2();
    px.l[0] ^= seed + threadIdx.x;
    py.l[1] ^= blockIdx.x;
    for (int i = 0; i < iters; i++) {
        acc.add_affine(px, py);
        px.l[2] += 1;
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) r ^= (uint64_t)(acc.x.l[i] ^ acc.y.l[i] ^ acc.zz.l[i] ^ acc.zzz.l[i]) << (i & 31);
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__device__ uint32_t g_mb_mod[12] = {0x00000001u, 0x8508c000u, 0x30000000u, 0x170b5d44u, 0xba094800u, 0x1ef3622fu,
                                    0x00f5138fu, 0x1a22d9f3u, 0x6ca1493bu, 0xc63b05c0u, 0x17c510eau, 0x01ae3a46u};
// kind 12: Fq product with the modulus in registers (opaque to ptxas)
__global__ void k_mb_mul_regmod(uint64_t* out, int iters, uint32_t seed) {
    Fq x = Fq::one(), y = Fq::r2();
    x.l[0] ^= seed + threadIdx.x;
    y.l[1] ^= blockIdx.x;
    uint32_t m[12];
#pragma unroll
    for (int i = 0; i < 12; i++) m[i] = ((volatile uint32_t*)g_mb_mod)[i + g_mb_lane_zero[threadIdx.x & 31]];
    for (int i = 0; i < iters; i++) {
        x = Fq::mul_m(x, y, m);
        y = Fq::mul_m(y, x, m);
    }
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) acc ^= (uint64_t)(x.l[i] ^ y.l[i]) << (i & 31);
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

cudaError_t microbench_run(int kind, int blocks, int threads, int iters, uint64_t* scratch, double* ops_per_launch,
                           cudaStream_t st) {
    double per_thread = 0;
    switch (kind) {
        case 0: k_mb_wide<<<blocks, threads, 0, st>>>(scratch, iters, 12345u); per_thread = 64.0 * iters; break;
        case 1: k_mb_pair<<<blocks, threads, 0, st>>>(scratch, iters, 12345u); per_thread = 64.0 * iters; break;
        case 2: k_mb_mul<Fr><<<blocks, threads, 0, st>>>(scratch, iters, 12345u); per_thread = 2.0 * iters; break;
        case 3: k_mb_mul<Fq><<<blocks, threads, 0, st>>>(scratch, iters, 12345u); per_thread = 2.0 * iters; break;
        case 4: k_mb_madd<<<blocks, threads, 0, st>>>(scratch, iters, 12345u); per_thread = 1.0 * iters; break;
        case 5: k_mb_wide_carry<<<blocks, threads, 0, st>>>(scratch, iters, 12345u); per_thread = 24.0 * iters; break;
        case 6: k_mb_addc<<<blocks, threads, 0, st>>>(scratch, iters, 12345u); per_thread = 96.0 * iters; break;
        case 7: k_mb_imad<<<blocks, threads, 0, st>>>(scratch, iters, 12345u); per_thread = 64.0 * iters; break;
        case 12: k_mb_mul_regmod<<<blocks, threads, 0, st>>>(scratch, iters, 12345u); per_thread = 2.0 * iters; break;
        case 13: k_mb_wide_carry_clean<<<blocks, threads, 0, st>>>(scratch, iters, 12345u); per_thread = 24.0 * iters; break;
        case 14: k_mb_wide_plus_add<<<blocks, threads, 0, st>>>(scratch, iters, 12345u); per_thread = 16.0 * iters; break;
        case 15: k_mb_mul_x2<<<blocks, threads, 0, st>>>(scratch, iters, 12345u); per_thread = 4.0 * iters; break;
        default: return cudaErrorInvalidValue;
    }
    CZK_LAUNCHED();
    *ops_per_launch = per_thread * (double)blocks * (double)threads;
    return cudaGetLastError();
}

}  // namespace czk
