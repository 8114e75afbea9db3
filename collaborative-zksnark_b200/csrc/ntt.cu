// Device kernels of the Fr NTT (see ntt.cuh for the design and the reference lines replaced).
#include "ntt.cuh"

#include "launch_count.hpp"

namespace czk {

__device__ __forceinline__ Fr ld_fr(const uint32_t* p, size_t i) {
    const uint4* q = reinterpret_cast<const uint4*>(p) + 2 * i;
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ Fr ldg_fr(const uint32_t* p, size_t i) {
    const uint4* q = reinterpret_cast<const uint4*>(p) + 2 * i;
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fr(uint32_t* p, size_t i, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p) + 2 * i;
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ Fr fr_from_u64x4(const uint64_t* c) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        r.l[2 * i] = (uint32_t)c[i];
        r.l[2 * i + 1] = (uint32_t)(c[i] >> 32);
    }
    return r;
}

struct Fr4 {
    uint64_t v[4];
};

// table[k] = c * base^k.  Each thread owns CHUNK consecutive entries: one pow, then a running product.
constexpr int POW_CHUNK = 32;
__global__ void k_powers(uint32_t* table, Fr4 base4, Fr4 c4, size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * POW_CHUNK;
    if (lo >= n) return;
    Fr base = fr_from_u64x4(base4.v);
    Fr cur = Fr::mul(fr_from_u64x4(c4.v), Fr::pow_u64(base, lo));
    size_t hi = lo + POW_CHUNK < n ? lo + POW_CHUNK : n;
    for (size_t k = lo; k < hi; k++) {
        st_fr(table, k, cur);
        cur = Fr::mul(cur, base);
    }
}

cudaError_t ntt_build_powers(uint32_t* table, const uint64_t base[4], const uint64_t c[4], size_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    Fr4 b, cc;
    for (int i = 0; i < 4; i++) {
        b.v[i] = base[i];
        cc.v[i] = c[i];
    }
    size_t threads = (n + POW_CHUNK - 1) / POW_CHUNK;
    unsigned blocks = (unsigned)((threads + 127) / 128);
    k_powers<<<blocks, 128, 0, st>>>(table, b, cc, n); CZK_LAUNCHED();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// One pass = r consecutive DIF stages on 2^r x 2^cl tiles held in shared memory (limb-major).
__global__ void __launch_bounds__(NTT_THREADS)
k_ntt_pass(uint32_t* __restrict__ data, const uint32_t* __restrict__ tw, int log_d, int s, int r, int cl, int inverse) {
    extern __shared__ uint32_t sm[];
    const int tile_log = r + cl;
    const int tile = 1 << tile_log;
    const int L = log_d - s - r;
    const size_t half_d = (size_t)1 << (log_d - 1);
    const size_t blk = blockIdx.x;
    const size_t lowblk = blk & (((size_t)1 << (L - cl)) - 1);
    const size_t hi = blk >> (L - cl);
    const size_t base = (hi << (r + L)) | (lowblk << cl);
    const int cmask = (1 << cl) - 1;

    for (int e = threadIdx.x; e < tile; e += NTT_THREADS) {
        int mid = e >> cl, lowc = e & cmask;
        size_t g = base | ((size_t)mid << L) | (size_t)lowc;
        Fr v = ld_fr(data, g);
#pragma unroll
        for (int k = 0; k < 8; k++) sm[k * tile + e] = v.l[k];
    }
    __syncthreads();

    for (int u = 0; u < r; u++) {
        const int b = r - 1 - u;
        const int t_stage = s + u;
        for (int p = threadIdx.x; p < (tile >> 1); p += NTT_THREADS) {
            int lowc = p & cmask;
            int q = p >> cl;
            int mid_lo = ((q >> b) << (b + 1)) | (q & ((1 << b) - 1));
            int e0 = (mid_lo << cl) | lowc;
            int e1 = e0 | (1 << (b + cl));
            size_t j = ((size_t)(mid_lo & ((1 << b) - 1)) << L) | (lowblk << cl) | (size_t)lowc;
            size_t ex = j << t_stage;
            Fr w;
            if (!inverse) {
                w = ldg_fr(tw, ex);
            } else {
                // w^-ex = -w^(D/2 - ex) for ex > 0
                w = (ex == 0) ? Fr::one() : Fr::neg(ldg_fr(tw, half_d - ex));
            }
            Fr a, c;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                a.l[k] = sm[k * tile + e0];
                c.l[k] = sm[k * tile + e1];
            }
            Fr sum = Fr::add(a, c);
            Fr diff = Fr::mul(Fr::sub(a, c), w);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                sm[k * tile + e0] = sum.l[k];
                sm[k * tile + e1] = diff.l[k];
            }
        }
        __syncthreads();
    }

    for (int e = threadIdx.x; e < tile; e += NTT_THREADS) {
        int mid = e >> cl, lowc = e & cmask;
        size_t g = base | ((size_t)mid << L) | (size_t)lowc;
        Fr v;
#pragma unroll
        for (int k = 0; k < 8; k++) v.l[k] = sm[k * tile + e];
        st_fr(data, g, v);
    }
}

cudaError_t ntt_run_passes(uint32_t* data, const uint32_t* tw, int log_d, bool inverse, cudaStream_t st) {
    if (log_d == 0) return cudaSuccess;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_ntt_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 << NTT_TILE_LOG);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    NttPlan plan = ntt_make_plan(log_d);
    for (int i = 0; i < plan.npass; i++) {
        const NttPlanPass& p = plan.pass[i];
        int tile_log = p.r + p.cl;
        size_t blocks = (size_t)1 << (log_d - tile_log);
        size_t smem = (size_t)32 << tile_log;
        k_ntt_pass<<<(unsigned)blocks, NTT_THREADS, smem, st>>>(data, tw, log_d, p.s, p.r, p.cl, inverse ? 1 : 0); CZK_LAUNCHED();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------
__global__ void k_scale_by_powers(uint32_t* data, const uint32_t* lo, const uint32_t* hi, int lo_log, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr f = Fr::mul(ldg_fr(lo, i & (((size_t)1 << lo_log) - 1)), ldg_fr(hi, i >> lo_log));
    st_fr(data, i, Fr::mul(ld_fr(data, i), f));
}

cudaError_t ntt_scale_by_powers(uint32_t* data, const uint32_t* lo, const uint32_t* hi, int lo_log, int log_d, cudaStream_t st) {
    size_t n = (size_t)1 << log_d;
    k_scale_by_powers<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(data, lo, hi, lo_log, n); CZK_LAUNCHED();
    return cudaGetLastError();
}

__global__ void k_bitrev_scale(uint32_t* data, int log_d, int mode, Fr4 c4, const uint32_t* lo, const uint32_t* hi, int lo_log) {
    size_t n = (size_t)1 << log_d;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t r = log_d == 0 ? 0 : (size_t)(__brevll((unsigned long long)i) >> (64 - log_d));
    if (i > r) return;
    Fr c = fr_from_u64x4(c4.v);
    Fr a = ld_fr(data, i);
    if (i == r) {
        if (mode == 1) a = Fr::mul(a, c);
        else if (mode == 2) a = Fr::mul(a, Fr::mul(ldg_fr(lo, i & (((size_t)1 << lo_log) - 1)), ldg_fr(hi, i >> lo_log)));
        if (mode) st_fr(data, i, a);
        return;
    }
    Fr b = ld_fr(data, r);
    // element from position r lands at i and vice versa; the factor follows the destination index
    if (mode == 1) {
        a = Fr::mul(a, c);
        b = Fr::mul(b, c);
    } else if (mode == 2) {
        Fr fi = Fr::mul(ldg_fr(lo, i & (((size_t)1 << lo_log) - 1)), ldg_fr(hi, i >> lo_log));
        Fr fr_ = Fr::mul(ldg_fr(lo, r & (((size_t)1 << lo_log) - 1)), ldg_fr(hi, r >> lo_log));
        b = Fr::mul(b, fi);
        a = Fr::mul(a, fr_);
    }
    st_fr(data, i, b);
    st_fr(data, r, a);
}

cudaError_t ntt_bitrev_scale(uint32_t* data, int log_d, int mode, const uint64_t c[4], const uint32_t* lo,
                             const uint32_t* hi, int lo_log, cudaStream_t st) {
    size_t n = (size_t)1 << log_d;
    Fr4 cc{};
    if (c)
        for (int i = 0; i < 4; i++) cc.v[i] = c[i];
    k_bitrev_scale<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(data, log_d, mode, cc, lo, hi, lo_log); CZK_LAUNCHED();
    return cudaGetLastError();
}

}  // namespace czk
