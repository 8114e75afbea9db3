// Fixed-base multi-scalar multiplication on the device: out[i] = s_i * B for one base B and n scalars - the CRS generator's
// workhorse.
//
// Replaces FixedBaseMSM::{get_window_table, windowed_mul, multi_scalar_mul} + batch_normalization_into_affine
// (algebra/ec/src/msm/fixed_base.rs:12-96, algebra/ec/src/models/short_weierstrass_jacobian.rs:480-500) as groth16's
// generator uses them (groth16/src/generator.rs:164-221).  Same function - every output is the affine point s_i * B, or
// infinity for s_i = 0 - computed with a fixed 8-bit window (32 windows of 255 table entries, ~0.8 MB for G1: L1/L2
// resident) instead of the reference's size-dependent window: one thread per scalar performs <= 32 mixed additions in XYZZ
// coordinates, then a second kernel normalises with Montgomery's trick (8 points per inversion).
#include "../../include/czk_groth16.h"
#include "ctx.hpp"
#include "launch_count.hpp"
#include "msm_io.cuh"

namespace czk {

constexpr int FB_WINDOW = 8, FB_WINDOWS = 32, FB_ENTRIES = 1 << FB_WINDOW;  // 32 * 8 = 256 >= 253 bits

template <class F>
__device__ __forceinline__ F fb_inverse(const F& a);
template <>
__device__ __forceinline__ Fq fb_inverse<Fq>(const Fq& a) {
    return Fq::inv_fermat(a);
}
template <>
__device__ __forceinline__ Fq2 fb_inverse<Fq2>(const Fq2& a) {
    return Fq2::inv_fermat(a);
}

// table[j * 256 + d] = d * 2^(8 j) * B as an affine point, d = 1 .. 255 (entry d = 0 is unused)
template <class F>
__global__ void __launch_bounds__(64) k_fb_table(uint32_t* __restrict__ table, const uint32_t* __restrict__ base_xy) {
    constexpr int W = FieldIO<F>::W;
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= FB_WINDOWS * FB_ENTRIES) return;
    const unsigned j = id / FB_ENTRIES, d = id % FB_ENTRIES;
    if (d == 0) return;
    const F bx = FieldIO<F>::load(base_xy), by = FieldIO<F>::load(base_xy + W);
    // d * B by double-and-add, then 8 j doublings
    XYZZ<F> acc = XYZZ<F>::infinity();
    for (int bit = FB_WINDOW - 1; bit >= 0; bit--) {
        acc = XYZZ<F>::dbl(acc);
        if ((d >> bit) & 1) acc.add_affine(bx, by);
    }
    for (unsigned k = 0; k < FB_WINDOW * j; k++) acc = XYZZ<F>::dbl(acc);
    // B has prime order r > 2^252 and d 2^(8j) < 2^256 is not a multiple of r for these (d, j): acc is finite unless B is not
    F inv = fb_inverse<F>(F::mul(acc.zz, acc.zzz));
    F ox = F::mul(acc.x, F::mul(inv, acc.zzz)), oy = F::mul(acc.y, F::mul(inv, acc.zz));
    FieldIO<F>::store(table + (size_t)id * (2 * W), ox);
    FieldIO<F>::store(table + (size_t)id * (2 * W) + W, oy);
}

// tmp[i] = s_i * B in XYZZ: one mixed addition per non-zero byte of the canonical scalar
template <class F>
__global__ void __launch_bounds__(128) k_fb_mul(uint32_t* __restrict__ tmp, const uint32_t* __restrict__ table,
                                                 const uint32_t* __restrict__ scalars, size_t n) {
    constexpr int W = FieldIO<F>::W;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* q = reinterpret_cast<const uint4*>(scalars) + 2 * i;
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fr s;
    s.l[0] = a.x; s.l[1] = a.y; s.l[2] = a.z; s.l[3] = a.w;
    s.l[4] = b.x; s.l[5] = b.y; s.l[6] = b.z; s.l[7] = b.w;
    s = Fr::from_mont(s);
    XYZZ<F> acc = XYZZ<F>::infinity();
#pragma unroll 1
    for (int j = 0; j < FB_WINDOWS; j++) {
        const unsigned d = (s.l[j >> 2] >> (8 * (j & 3))) & 0xffu;
        if (d) {
            const uint32_t* p = table + ((size_t)j * FB_ENTRIES + d) * (2 * W);
            acc.add_affine(FieldIO<F>::load(p), FieldIO<F>::load(p + W));
        }
    }
    store_point<F>(tmp + i * (4 * W), acc);
}

// XYZZ -> affine with one inversion per FB_NORM points (Montgomery's trick); infinity -> (0, 1) + flag
constexpr int FB_NORM = 8;
template <class F>
__global__ void __launch_bounds__(128) k_fb_normalize(uint32_t* __restrict__ out_xy, uint8_t* __restrict__ out_inf,
                                                       const uint32_t* __restrict__ tmp, size_t n) {
    constexpr int W = FieldIO<F>::W;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    F pre[FB_NORM];
    F acc = F::one();
#pragma unroll
    for (int j = 0; j < FB_NORM; j++) {
        const size_t i = t + (size_t)j * stride;
        F zz = F::one(), zzz = F::one();
        if (i < n) {
            zz = FieldIO<F>::load_rw(tmp + i * (4 * W) + 2 * W);
            zzz = FieldIO<F>::load_rw(tmp + i * (4 * W) + 3 * W);
            if (zz.is_zero()) zz = zzz = F::one();
        }
        pre[j] = acc;
        acc = F::mul(acc, F::mul(zz, zzz));
    }
    F inv = fb_inverse<F>(acc);
#pragma unroll
    for (int j = FB_NORM - 1; j >= 0; j--) {
        const size_t i = t + (size_t)j * stride;
        if (i >= n) continue;
        const uint32_t* p = tmp + i * (4 * W);
        F zz = FieldIO<F>::load_rw(p + 2 * W), zzz = FieldIO<F>::load_rw(p + 3 * W);
        const bool inf = zz.is_zero();
        if (inf) zz = zzz = F::one();
        F ti = F::mul(inv, pre[j]);  // 1 / (zz zzz)
        inv = F::mul(inv, F::mul(zz, zzz));
        F ox = F::zero(), oy = F::one();
        if (!inf) {
            ox = F::mul(FieldIO<F>::load_rw(p), F::mul(ti, zzz));      // X / zz
            oy = F::mul(FieldIO<F>::load_rw(p + W), F::mul(ti, zz));   // Y / zzz
        }
        FieldIO<F>::store(out_xy + i * (2 * W), ox);
        FieldIO<F>::store(out_xy + i * (2 * W) + W, oy);
        out_inf[i] = inf ? 1 : 0;
    }
}

template <class F>
static cudaError_t fixed_base_t(uint32_t* out_xy, uint8_t* out_inf, uint32_t* table, uint32_t* tmp, const uint32_t* base_xy,
                                const uint32_t* scalars, size_t n, cudaStream_t st) {
    k_fb_table<F><<<(FB_WINDOWS * FB_ENTRIES + 63) / 64, 64, 0, st>>>(table, base_xy); CZK_LAUNCHED();
    if (n) {
        k_fb_mul<F><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(tmp, table, scalars, n); CZK_LAUNCHED();
        size_t threads = (n + FB_NORM - 1) / FB_NORM;
        k_fb_normalize<F><<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(out_xy, out_inf, tmp, n); CZK_LAUNCHED();
    }
    return cudaGetLastError();
}

}  // namespace czk

using namespace czk;

// bases[i] = scalars[sc_off + i] * base, i < n, as a resident base set (infinity flags where the scalar is zero)
int czk_fixed_base_msm(czk_ctx* ctx, int curve, const uint64_t* base_xy, const czk_vec* scalars, size_t sc_off, size_t n,
                       czk_bases** out) {
    if (!ctx || !out || !base_xy || (curve != 1 && curve != 2) || (n && (!scalars || sc_off + n > scalars->n)))
        return fail(ctx, CZK_ERR_ARG, "czk_fixed_base_msm: argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t pb = curve == 1 ? 96 : 192;
    czk_bases* b = new czk_bases();
    b->curve = curve;
    b->n = n;
    CUDA_TRY(ctx, cudaMalloc((void**)&b->xy, (n ? n : 1) * pb));
    CUDA_TRY(ctx, cudaMalloc((void**)&b->inf, n ? n : 1));
    uint32_t *table = nullptr, *tmp = nullptr, *base_dev = nullptr;
    CUDA_TRY(ctx, cudaMallocAsync((void**)&table, (size_t)FB_WINDOWS * FB_ENTRIES * pb, ctx->stream));
    CUDA_TRY(ctx, cudaMallocAsync((void**)&tmp, (n ? n : 1) * 2 * pb, ctx->stream));
    CUDA_TRY(ctx, cudaMallocAsync((void**)&base_dev, pb, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(base_dev, base_xy, pb, cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t* sc = n ? (const uint32_t*)(scalars->d + 4 * sc_off) : nullptr;
    cudaError_t e = curve == 1 ? fixed_base_t<Fq>(b->xy, b->inf, table, tmp, base_dev, sc, n, ctx->stream)
                               : fixed_base_t<Fq2>(b->xy, b->inf, table, tmp, base_dev, sc, n, ctx->stream);
    CUDA_TRY(ctx, e);
    CUDA_TRY(ctx, cudaFreeAsync(table, ctx->stream));
    CUDA_TRY(ctx, cudaFreeAsync(tmp, ctx->stream));
    CUDA_TRY(ctx, cudaFreeAsync(base_dev, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *out = b;
    return CZK_OK;
}
