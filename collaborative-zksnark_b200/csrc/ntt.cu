// Device kernels of the Fr NTT (see ntt.cuh for the design and the reference lines replaced).
#include "ntt.cuh"
#include "ntt_tile.cuh"

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
// One pass on one tile per block (ntt_tile.cuh has the per-thread arithmetic).  The tile travels by bulk asynchronous
// copies (TMA, cp.async.bulk): warp 0 issues one copy per tile row into shared memory, every thread waits on the
// mbarrier the copies complete on, the register phases run with one __syncthreads() between them, and warp 0 sends the
// rows back with bulk stores.  No thread ever computes a global address for data, and the tile's layout in shared
// memory is the vector's own (32 bytes per element), so a row is one contiguous copy.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}

struct NttVecs {
    uint32_t* p[NTT_MAX_BATCH];
};
struct SmemTile {
    uint4* t;  // element e = t[2e], t[2e + 1]
    __device__ __forceinline__ Fr load(unsigned e) const {
        const uint4 a = t[2 * e], b = t[2 * e + 1];
        Fr r;
        r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
        r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
        return r;
    }
    __device__ __forceinline__ void store(unsigned e, const Fr& v) const {
        t[2 * e] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        t[2 * e + 1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    }
};

template <bool DIT, bool SCALE>
__global__ void __launch_bounds__(NTT_TILE_THREADS, NTT_BLOCKS_PER_SM)
k_ntt_tile(NttVecs vecs, const uint32_t* __restrict__ tw, int n, int s, int r, int cl, int inverse, int first_pass, int last_pass,
           NttScale pre, NttScale post) {
    extern __shared__ __align__(128) uint4 tile4[];
    __shared__ __align__(8) uint64_t bar;
    const int tile_log = r + cl, L = n - s - r;
    const unsigned te = 1u << tile_log, tid = threadIdx.x;
    uint32_t* const data = vecs.p[blockIdx.y];
    const size_t blk = blockIdx.x;
    const size_t lowblk = blk & (((size_t)1 << (L - cl)) - 1), hi = blk >> (L - cl);
    NttTileGeom g{n, s, r, cl, L, (hi << (r + L)) | (lowblk << cl)};
    // rows of the tile in the vector: 2^r rows of 2^cl elements, 2^L apart - one contiguous run when L == cl
    const unsigned rows = L == cl ? 1u : (1u << r);
    const unsigned row_bytes = (L == cl ? te : (1u << cl)) * 32u;
    const size_t row_stride = (size_t)1 << L;  // elements
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid < 32) {
        if (tid == 0) mbar_expect_tx(&bar, te * 32u);
        __syncwarp();
        for (unsigned m = tid; m < rows; m += 32)
            bulk_g2s(reinterpret_cast<uint8_t*>(tile4) + (size_t)m * row_bytes, data + (g.base + m * row_stride) * 8, row_bytes, &bar);
    }
    SmemTile tile{tile4};
    auto ldtw = [](const uint32_t* table, size_t i) -> Fr { return ldg_fr(table, i); };
    const int nph = ntt_num_phases(r);
    mbar_wait(&bar, 0);
    for (int ph = 0; ph < nph; ph++) {
        int kp, jlo, ns;
        ntt_phase_geom(r, cl, DIT, ph, kp, jlo, ns);
        if (tid < te / 8)
            ntt_phase_thread<DIT, SCALE>(tile, tid, g, inverse != 0, kp, jlo, ns, tw, first_pass && ph == 0, last_pass && ph == nph - 1, pre, post,
                                         ldtw);
        if (ph + 1 < nph) __syncthreads();
    }
    // the generic-proxy writes of every thread must be visible to the bulk-copy (async proxy) reads
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid < 32) {
        for (unsigned m = tid; m < rows; m += 32)
            bulk_s2g(data + (g.base + m * row_stride) * 8, reinterpret_cast<uint8_t*>(tile4) + (size_t)m * row_bytes, row_bytes);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory may be released once it has been read
    }
}

// log_d <= 2: the whole transform in one thread (naive DFT over at most 4 points)
__global__ void k_ntt_small(NttVecs vecs, const uint32_t* __restrict__ tw, int n, int inverse, NttScale pre, NttScale post) {
    if (threadIdx.x || blockIdx.x) return;
    uint32_t* data = vecs.p[blockIdx.y];
    const unsigned d = 1u << n, half = d >> 1;
    auto ldtw = [](const uint32_t* table, size_t i) -> Fr { return ldg_fr(table, i); };
    Fr x[4], y[4];
    for (unsigned i = 0; i < d; i++) {
        x[i] = ld_fr(data, i);
        if (pre.mode) x[i] = Fr::mul(x[i], ntt_scale_factor(pre, i, n, ldtw));
    }
    for (unsigned i = 0; i < d; i++) {
        Fr acc = Fr::zero();
        for (unsigned j = 0; j < d; j++) {
            unsigned e = (i * j) & (d - 1);
            if (inverse) e = (d - e) & (d - 1);
            Fr w = e < half || half == 0 ? ldg_fr(tw, e) : Fr::neg(ldg_fr(tw, e - half));  // omega^(e) = -omega^(e - d/2)
            acc = Fr::add(acc, Fr::mul(x[j], w));
        }
        y[i] = acc;
    }
    for (unsigned i = 0; i < d; i++) {
        Fr v = y[i];
        if (post.mode) v = Fr::mul(v, ntt_scale_factor(post, i, n, ldtw));
        st_fr(data, i, v);
    }
}

// All the passes of one in-place transform over `count` vectors.  dit = false: natural order in, bit-reversed out;
// dit = true: bit-reversed in, natural order out.  (log_d <= 2: natural in and out whatever `dit` says.)
cudaError_t ntt_run_tiles(uint32_t* const* data, int count, const uint32_t* tw, int log_d, bool inverse, bool dit, const NttScale& pre,
                          const NttScale& post, cudaStream_t st) {
    if (log_d == 0 || count == 0) return cudaSuccess;
    if (count > NTT_MAX_BATCH) return cudaErrorInvalidValue;
    NttVecs v{};
    for (int i = 0; i < count; i++) v.p[i] = data[i];
    if (log_d <= 2) {
        k_ntt_small<<<dim3(1, count), 32, 0, st>>>(v, tw, log_d, inverse ? 1 : 0, pre, post); CZK_LAUNCHED();
        return cudaGetLastError();
    }
    static const cudaError_t attr = [] {
        cudaError_t e = cudaFuncSetAttribute(k_ntt_tile<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 << NTT_TILE_LOG);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ntt_tile<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 << NTT_TILE_LOG);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ntt_tile<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 << NTT_TILE_LOG);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ntt_tile<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 << NTT_TILE_LOG);
        return e;
    }();
    if (attr != cudaSuccess) return attr;
    const NttPlan plan = ntt_make_plan(log_d);
    for (int k = 0; k < plan.npass; k++) {
        const NttPass& p = plan.pass[dit ? plan.npass - 1 - k : k];
        const int tile_log = p.r + p.cl;
        const unsigned blocks = 1u << (log_d - tile_log);
        const size_t smem = (size_t)32 << tile_log;
        const unsigned threads = (1u << tile_log) / 8 < 32 ? 32 : (1u << tile_log) / 8;
        const int first = k == 0, last = k == plan.npass - 1;
        const bool scale = (first && pre.mode) || (last && post.mode);  // only those passes run the variant that carries scaling code
        const dim3 grid(blocks, count);
        if (dit && scale) k_ntt_tile<true, true><<<grid, threads, smem, st>>>(v, tw, log_d, p.s, p.r, p.cl, inverse ? 1 : 0, first, last, pre, post);
        else if (dit) k_ntt_tile<true, false><<<grid, threads, smem, st>>>(v, tw, log_d, p.s, p.r, p.cl, inverse ? 1 : 0, first, last, pre, post);
        else if (scale) k_ntt_tile<false, true><<<grid, threads, smem, st>>>(v, tw, log_d, p.s, p.r, p.cl, inverse ? 1 : 0, first, last, pre, post);
        else k_ntt_tile<false, false><<<grid, threads, smem, st>>>(v, tw, log_d, p.s, p.r, p.cl, inverse ? 1 : 0, first, last, pre, post);
        CZK_LAUNCHED();
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

// --------------------------------------------------------------------------------------------- mixed radix (3 * 2^k)
// MixedRadixEvaluationDomain (algebra/poly/src/domain/mixed_radix.rs) for Fr's small subgroup base 3: a transform over
// N = 3 M points (M = 2^k) with w = get_root_of_unity(N).  Writing the input index as 3 a + r,
//   X[j + s M] = Y_0[j] + zeta^s w^j Y_1[j] + zeta^(2 s) w^(2 j) Y_2[j],   Y_r = DFT_M of x[3 a + r] under w^3,
// with zeta = w^M a primitive cube root of unity and w^3 the radix-2 domain's own generator: de-interleave, three radix-2
// transforms of M points (one batched grid of the tile kernels above), one combining pass.  The reference's serial
// permute + radix-3 + radix-2 passes compute the same DFT; nothing of their order is kept.
__global__ void k_mr_split(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t M) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * M) return;
    st_fr(out, (i % 3) * M + i / 3, ld_fr(in, i));
}
// wpow[j] = w^j (or w^-j with zeta^-1 for the inverse), j < M; scale: multiply everything by c (3^-1 on the way back)
__global__ void k_mr_combine(const uint32_t* __restrict__ y, uint32_t* __restrict__ out, const uint32_t* __restrict__ wpow, Fr4 zeta4,
                             Fr4 c4, int scale, size_t M) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const Fr yv[3] = {ld_fr(y, j), ld_fr(y, M + j), ld_fr(y, 2 * M + j)};
    Fr out3[3];
    ntt_mixed_combine3(yv, ldg_fr(wpow, j), fr_from_u64x4(zeta4.v), fr_from_u64x4(c4.v), scale != 0, out3);
    st_fr(out, j, out3[0]);
    st_fr(out, M + j, out3[1]);
    st_fr(out, 2 * M + j, out3[2]);
}
cudaError_t ntt_mixed_split(const uint32_t* in, uint32_t* out, size_t M, cudaStream_t st) {
    k_mr_split<<<(unsigned)((3 * M + 255) / 256), 256, 0, st>>>(in, out, M); CZK_LAUNCHED();
    return cudaGetLastError();
}
cudaError_t ntt_mixed_combine(const uint32_t* y, uint32_t* out, const uint32_t* wpow, const uint64_t zeta[4], const uint64_t c[4], size_t M,
                              cudaStream_t st) {
    Fr4 z{}, cc{};
    for (int i = 0; i < 4; i++) {
        z.v[i] = zeta[i];
        cc.v[i] = c ? c[i] : 0;
    }
    k_mr_combine<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(y, out, wpow, z, cc, c ? 1 : 0, M); CZK_LAUNCHED();
    return cudaGetLastError();
}

}  // namespace czk
