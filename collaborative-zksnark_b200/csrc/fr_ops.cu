#include "fr_ops.cuh"

#include "launch_count.hpp"

namespace czk {

__device__ __forceinline__ Fr ldv(const uint32_t* p, size_t i) {
    const uint4* q = reinterpret_cast<const uint4*>(p) + 2 * i;
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void stv(uint32_t* p, size_t i, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p) + 2 * i;
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ Fr cst(const FrConst& c) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        r.l[2 * i] = (uint32_t)c.v[i];
        r.l[2 * i + 1] = (uint32_t)(c.v[i] >> 32);
    }
    return r;
}
static FrConst mk(const uint64_t c[4]) {
    FrConst r;
    for (int i = 0; i < 4; i++) r.v[i] = c[i];
    return r;
}
static unsigned grid_for(size_t n, int threads) {
    size_t b = (n + threads - 1) / threads;
    size_t cap = 148 * 16;  // grid-stride: a few waves of the 148 SMs
    return (unsigned)(b < cap ? (b ? b : 1) : cap);
}
#define GRID_STRIDE(i, n) for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (size_t)gridDim.x * blockDim.x)

__global__ void k_fr_binop(uint32_t* a, const uint32_t* b, size_t n, int op) {
    GRID_STRIDE(i, n) {
        Fr x = ldv(a, i), y = ldv(b, i);
        Fr r = op == FR_ADD ? Fr::add(x, y) : op == FR_SUB ? Fr::sub(x, y) : Fr::mul(x, y);
        stv(a, i, r);
    }
}
cudaError_t fr_binop(uint32_t* a, const uint32_t* b, size_t n, FrBinOp op, cudaStream_t st) {
    if (!n) return cudaSuccess;
    k_fr_binop<<<grid_for(n, 256), 256, 0, st>>>(a, b, n, (int)op); CZK_LAUNCHED();
    return cudaGetLastError();
}

__global__ void k_fr_scale(uint32_t* a, FrConst c, size_t n) {
    Fr cc = cst(c);
    GRID_STRIDE(i, n) stv(a, i, Fr::mul(ldv(a, i), cc));
}
cudaError_t fr_scale(uint32_t* a, const uint64_t c[4], size_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    k_fr_scale<<<grid_for(n, 256), 256, 0, st>>>(a, mk(c), n); CZK_LAUNCHED();
    return cudaGetLastError();
}

__global__ void k_fr_scale_by_tables(uint32_t* a, const uint32_t* lo, const uint32_t* hi, int lo_log, size_t n) {
    GRID_STRIDE(i, n) {
        Fr f = Fr::mul(ldv(lo, i & (((size_t)1 << lo_log) - 1)), ldv(hi, i >> lo_log));
        stv(a, i, Fr::mul(ldv(a, i), f));
    }
}
cudaError_t fr_scale_by_tables(uint32_t* a, const uint32_t* lo, const uint32_t* hi, int lo_log, size_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    k_fr_scale_by_tables<<<grid_for(n, 256), 256, 0, st>>>(a, lo, hi, lo_log, n);
    CZK_LAUNCHED();
    return cudaGetLastError();
}

__global__ void k_fr_add_const(uint32_t* d, const uint32_t* x, FrConst t, size_t n) {
    Fr tt = cst(t);
    GRID_STRIDE(i, n) stv(d, i, Fr::add(ldv(x, i), tt));
}
cudaError_t fr_add_const(uint32_t* d, const uint32_t* x, const uint64_t tx[4], size_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    k_fr_add_const<<<grid_for(n, 256), 256, 0, st>>>(d, x, mk(tx), n); CZK_LAUNCHED();
    return cudaGetLastError();
}

}  // namespace czk
