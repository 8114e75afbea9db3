// Integer-pipe microbenchmarks: the roofline denominators for the field kernels.  MEASURED_PEAKS.json
// carries only HBM and bf16 peaks; these measure IMAD / IMAD.WIDE issue rates and the achieved field
// multiplication and point addition rates at a chosen occupancy.
#include <cuda_runtime.h>
#include "ec.cuh"
#include "fq13.cuh"
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
// kind 8: independent mad.hi chains
__global__ void k_mb_imad_hi(uint64_t* out, int iters, uint32_t seed) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    uint32_t c0 = 1, c1 = 2, c2 = 3, c3 = 4, c4 = 5, c5 = 6, c6 = 7, c7 = 8;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(c0) : "r"(a), "r"(b));
            asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(c1) : "r"(a), "r"(b));
            asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(c2) : "r"(a), "r"(b));
            asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(c3) : "r"(a), "r"(b));
            asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(c4) : "r"(a), "r"(b));
            asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(c5) : "r"(a), "r"(b));
            asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(c6) : "r"(a), "r"(b));
            asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(c7) : "r"(a), "r"(b));
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

__device__ uint32_t g_mb_zero = 0;
__global__ void k_mb_mul_split(uint64_t* out, int iters, uint32_t seed) {
    Fq x = Fq::one(), y = Fq::r2();
    x.l[0] ^= seed + threadIdx.x;
    y.l[1] ^= blockIdx.x;
    const uint32_t z = *(volatile uint32_t*)&g_mb_zero;
    for (int i = 0; i < iters; i++) {
        x = Fq::mul_split(x, y, z);
        y = Fq::mul_split(y, x, z);
    }
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) acc ^= (uint64_t)(x.l[i] ^ y.l[i]) << (i & 31);
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_mb_mul13(uint64_t* out, int iters, uint32_t seed) {
    Fq13 x = Fq13::one(), y = Fq13::one();
    x.d[0] ^= (seed + threadIdx.x) & 0xffff;
    y.d[1] ^= blockIdx.x & 0xffff;
    for (int i = 0; i < iters; i++) {
        x = Fq13::mul(x, y);
        y = Fq13::mul(y, x);
    }
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 13; i++) acc ^= (uint64_t)(x.d[i] ^ y.d[i]) << (i & 31);
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_mb_madd13(uint64_t* out, int iters, uint32_t seed) {
    XYZZ<Fq13> acc = XYZZ<Fq13>::from_affine(Fq13::one(), Fq13::one());
    Fq13 px = Fq13::one(), py = Fq13::one();
    px.d[0] ^= (seed + threadIdx.x) & 0xffff;
    py.d[1] ^= blockIdx.x & 0xffff;
    for (int i = 0; i < iters; i++) {
        acc.add_affine(px, py);
        px.d[2] = (px.d[2] + 1) & Fq13::MASK;
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < 13; i++) r ^= (uint64_t)(acc.x.d[i] ^ acc.y.d[i] ^ acc.zz.d[i] ^ acc.zzz.d[i]) << (i & 31);
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
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
        case 8: k_mb_imad_hi<<<blocks, threads, 0, st>>>(scratch, iters, 12345u); per_thread = 64.0 * iters; break;
        case 11: k_mb_mul_split<<<blocks, threads, 0, st>>>(scratch, iters, 12345u); per_thread = 2.0 * iters; break;
        case 9: k_mb_mul13<<<blocks, threads, 0, st>>>(scratch, iters, 12345u); per_thread = 2.0 * iters; break;
        case 10: k_mb_madd13<<<blocks, threads, 0, st>>>(scratch, iters, 12345u); per_thread = 1.0 * iters; break;
        default: return cudaErrorInvalidValue;
    }
    CZK_LAUNCHED();
    *ops_per_launch = per_thread * (double)blocks * (double)threads;
    return cudaGetLastError();
}

}  // namespace czk
