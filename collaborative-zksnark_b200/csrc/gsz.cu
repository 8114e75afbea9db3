// GSZ20 honest-majority (Shamir) shares on the device: open / king_compute / batch_mult and the product checks.
//
// Replaces, for the Groth16 path, mpc-algebra/src/share/gsz20/mod.rs:
//   :94-105   t = (n-1)/2, share domain = MixedRadixEvaluationDomain(n): party j holds p(w^j)
//   :434-459  open / open_degree_vec: interpolate over the share domain, assert degree <= d, evaluate at 0
//   :468-524  king_compute / batch_king_compute: gather to the king, king opens, king returns the value to everyone
//   :529-594  coin, mult, batch_mult (local product + double-random mask, king degree reduction, triple queued)
//   :599-808  hadamard_check -> ip_check (ip_compute, ip_compress) run at the first reveal
// Every party's n x k share matrix lives in HBM; the interpolation is one streaming kernel over the gathered shares
// (n_inv * sum_j s_j for the value, sum_j s_j w^-jk == 0 for every k > d as the degree check), the exchanges are NCCL:
// grouped send/recv to rank 0 + broadcast for king_compute, all-gather for open.  The reference's preprocessing is
// stubbed (rand() = 1, double_rand() = (1, 1), the king re-shares the plain value): the same stubs are used here, so
// results are bit-identical; the interpolation and the degree check nevertheless run on the real gathered shares.
#include "ctx.hpp"
#include "fr_ops.cuh"
#include "launch_count.hpp"
#include "ntt.cuh"

namespace czk {

__device__ __forceinline__ Fr g_ld(const uint32_t* p, size_t i) {
    const uint4* q = reinterpret_cast<const uint4*>(p) + 2 * i;
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void g_st(uint32_t* p, size_t i, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p) + 2 * i;
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ Fr g_cst(const FrConst& c) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        r.l[2 * i] = (uint32_t)c.v[i];
        r.l[2 * i + 1] = (uint32_t)(c.v[i] >> 32);
    }
    return r;
}
static FrConst g_mk(const uint64_t c[4]) {
    FrConst r;
    for (int i = 0; i < 4; i++) r.v[i] = c[i];
    return r;
}
static unsigned g_grid(size_t n, int threads) {
    size_t b = (n + threads - 1) / threads;
    size_t cap = 148 * 8;
    return (unsigned)(b < cap ? (b ? b : 1) : cap);
}
#define G_STRIDE(i, n) for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (size_t)gridDim.x * blockDim.x)

// open_degree_vec over k elements: g = parties x k gathered shares (party-major).
//   out[i] = n^-1 sum_j g[j][i]                       (coefficient 0 of the inverse DFT = value at 0)
//   flag  |= any  sum_j g[j][i] w^(-j c) != 0, c > d   (coefficients above the degree bound must vanish)
__global__ void k_gsz_open(uint32_t* __restrict__ out, const uint32_t* __restrict__ g, size_t k, int parties, int degree,
                           const uint32_t* __restrict__ winv, FrConst n_inv, uint32_t* __restrict__ flag) {
    const Fr ninv = g_cst(n_inv);
    uint32_t bad = 0;
    G_STRIDE(i, k) {
        Fr acc = g_ld(g, i);
        for (int j = 1; j < parties; j++) acc = Fr::add(acc, g_ld(g, (size_t)j * k + i));
        g_st(out, i, Fr::mul(acc, ninv));
        for (int c = degree + 1; c < parties; c++) {
            Fr coef = g_ld(g, i);  // j = 0: w^0
            for (int j = 1; j < parties; j++) coef = Fr::add(coef, Fr::mul(g_ld(g, (size_t)j * k + i), g_ld(winv, (size_t)((j * c) % parties))));
            if (!coef.is_zero()) bad = 1;
        }
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

// d[i] = x[i] * y[i] + c      (the masked local product of batch_mult, :571-577)
__global__ void k_gsz_mul_add_const(uint32_t* __restrict__ d, const uint32_t* __restrict__ x, const uint32_t* __restrict__ y, FrConst c,
                                    size_t n) {
    const Fr cc = g_cst(c);
    G_STRIDE(i, n) g_st(d, i, Fr::add(Fr::mul(g_ld(x, i), g_ld(y, i)), cc));
}

// block partial sums of: mode 0  x[i]*y[i] ; mode 1  (2 xr - xl)(2 yr - yl) with xr = x + h, yr = y + h (the value at 3
// of the lines through (1, left), (2, right): ip_compress, :664-679) ; mode 2  x[i]
__global__ void __launch_bounds__(256) k_gsz_dot(uint32_t* __restrict__ partial, const uint32_t* __restrict__ x,
                                                  const uint32_t* __restrict__ y, size_t h, int mode) {
    __shared__ uint32_t sm[256 * 8];
    Fr acc = Fr::zero();
    G_STRIDE(i, h) {
        Fr t;
        if (mode == 0) t = Fr::mul(g_ld(x, i), g_ld(y, i));
        else if (mode == 1) {
            Fr xl = g_ld(x, i), xr = g_ld(x, h + i), yl = g_ld(y, i), yr = g_ld(y, h + i);
            Fr x3 = Fr::add(xr, Fr::sub(xr, xl)), y3 = Fr::add(yr, Fr::sub(yr, yl));
            t = Fr::mul(x3, y3);
        } else t = g_ld(x, i);
        acc = Fr::add(acc, t);
    }
#pragma unroll
    for (int w = 0; w < 8; w++) sm[w * 256 + threadIdx.x] = acc.l[w];
    __syncthreads();
    for (unsigned s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            Fr a, b;
#pragma unroll
            for (int w = 0; w < 8; w++) a.l[w] = sm[w * 256 + threadIdx.x], b.l[w] = sm[w * 256 + threadIdx.x + s];
            a = Fr::add(a, b);
#pragma unroll
            for (int w = 0; w < 8; w++) sm[w * 256 + threadIdx.x] = a.l[w];
        }
        __syncthreads();
    }
    if (threadIdx.x < 8) partial[(size_t)blockIdx.x * 8 + threadIdx.x] = sm[threadIdx.x * 256];
}
// out[0] = sum of `count` partials + c
__global__ void k_gsz_dot_finish(uint32_t* __restrict__ out, const uint32_t* __restrict__ partial, unsigned count, FrConst c) {
    if (threadIdx.x || blockIdx.x) return;
    Fr acc = g_cst(c);
    for (unsigned b = 0; b < count; b++) acc = Fr::add(acc, g_ld(partial, b));
    g_st(out, 0, acc);
}
// ip_check fold (:708-721): x[i] <- (xr - xl) r + (2 xl - xr), same for y, i < h
__global__ void k_gsz_fold(uint32_t* __restrict__ x, uint32_t* __restrict__ y, size_t h, FrConst r) {
    const Fr rr = g_cst(r);
    G_STRIDE(i, h) {
        Fr xl = g_ld(x, i), xr = g_ld(x, h + i), yl = g_ld(y, i), yr = g_ld(y, h + i);
        Fr xm = Fr::sub(xr, xl), ym = Fr::sub(yr, yl);
        g_st(x, i, Fr::add(Fr::mul(xm, rr), Fr::sub(xl, xm)));
        g_st(y, i, Fr::add(Fr::mul(ym, rr), Fr::sub(yl, ym)));
    }
}

}  // namespace czk

using namespace czk;

// ------------------------------------------------------------------------------------------ share domain
// FftField::get_root_of_unity(n) (algebra/ff/src/fields/mod.rs:337-386), large-subgroup branch: n = 2^a 3^b, b <= 1
static bool gsz_root_of_unity(size_t n, HFr* out) {
    unsigned a = 0, b = 0;
    size_t m = n;
    while (m % 2 == 0) m /= 2, a++;
    while (m % 3 == 0) m /= 3, b++;
    if (m != 1 || a > FrParams::TWO_ADICITY || b > 1) return false;
    HFr w = HFr::from_limbs(FrParams::LARGE_ROOT_64);
    for (unsigned i = b; i < 1; i++) w = HFr::pow_u64(w, 3);
    for (unsigned i = a; i < FrParams::TWO_ADICITY; i++) w = HFr::sqr(w);
    *out = w;
    return true;
}

static int gsz_prepare(czk_ctx* ctx) {
    GszState& g = ctx->gsz;
    if (g.n == ctx->nranks && g.winv_dev) return CZK_OK;
    const int n = ctx->nranks;
    if (n > 64) return fail(ctx, CZK_ERR_ARG, "GSZ: more than 64 parties");
    HFr w;
    if (!gsz_root_of_unity((size_t)n, &w)) return fail(ctx, CZK_ERR_ARG, "GSZ: no share domain of this size (n must be 2^a or 3 * 2^a)");
    HFr wi = HFr::inv(w);
    std::vector<uint64_t> tab((size_t)n * 4);
    HFr p = HFr::one();
    g.winv_host.assign((size_t)n, HFr::one());
    for (int k = 0; k < n; k++) {
        p.to_limbs(tab.data() + 4 * k);
        g.winv_host[(size_t)k] = p;
        p = HFr::mul(p, wi);
    }
    if (g.winv_dev) cudaFree(g.winv_dev);
    g.winv_dev = nullptr;
    CUDA_TRY(ctx, cudaMalloc((void**)&g.winv_dev, (size_t)n * 32));
    CUDA_TRY(ctx, cudaMemcpyAsync(g.winv_dev, tab.data(), (size_t)n * 32, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    g.n_inv = HFr::inv(HFr::from_u64((uint64_t)n));
    g.n = n;
    g.t = (n - 1) / 2;
    return CZK_OK;
}

static int gsz_check_flag(czk_ctx* ctx, const char* what) {
    uint32_t flag = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&flag, ctx->flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (flag) {
        cudaMemsetAsync(ctx->flag, 0, 4, ctx->stream);
        return fail(ctx, CZK_ERR_PROTOCOL, std::string(what) + ": degree check failed (gsz20/mod.rs:449 assert!(p.degree() <= d))");
    }
    return CZK_OK;
}

// send_bytes_to_king (mpc-net/src/multi.rs:176-209): rank 0 receives every party's `bytes` into recv[p * bytes]
int czk_net_gather_to_king_dev(czk_ctx* ctx, const void* dev_send, void* dev_recv_king, size_t bytes) {
    if (!ctx || !dev_send || (ctx->rank == 0 && !dev_recv_king)) return fail(ctx, CZK_ERR_ARG, "czk_net_gather_to_king_dev: null");
    if (ctx->rank == 0) ctx->stats[1] += bytes * (uint64_t)(ctx->nranks - 1);
    else ctx->stats[0] += bytes;
    ctx->stats[3] += 1;
    if (ctx->rank == 0 && dev_recv_king != dev_send)
        CUDA_TRY(ctx, cudaMemcpyAsync(dev_recv_king, dev_send, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    if (ctx->nranks == 1) return CZK_OK;
    NcclApi& api = nccl_api();
    ncclResult_t r = api.GroupStart();
    if (r == ncclSuccess) {
        if (ctx->rank == 0) {
            for (int p = 1; p < ctx->nranks && r == ncclSuccess; p++)
                r = api.Recv((uint8_t*)dev_recv_king + (size_t)p * bytes, bytes, ncclUint8, p, ctx->comm, ctx->stream);
        } else {
            r = api.Send(dev_send, bytes, ncclUint8, 0, ctx->comm, ctx->stream);
        }
    }
    ncclResult_t r2 = api.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) return fail(ctx, CZK_ERR_NCCL, std::string("gather to king: ") + api.GetErrorString(r));
    return CZK_OK;
}

// king_compute with f = identity on k device elements, in place: v <- the value the king opened at `degree`
static int gsz_king_compute_dev(czk_ctx* ctx, uint32_t* v, size_t k, int degree) {
    GszState& g = ctx->gsz;
    CZK_TRY(gsz_prepare(ctx));
    g.king_computes++;
    const size_t bytes = k * 32;
    if (ctx->rank == 0) CZK_TRY(scratch_reserve(ctx, g.gather, bytes * (size_t)ctx->nranks));
    CZK_TRY(czk_net_gather_to_king_dev(ctx, v, ctx->rank == 0 ? g.gather.p : nullptr, bytes));
    if (ctx->rank == 0 && k) {
        k_gsz_open<<<g_grid(k, 256), 256, 0, ctx->stream>>>(v, (const uint32_t*)g.gather.p, k, ctx->nranks, degree, g.winv_dev, g_mk(g.n_inv.l),
                                                            ctx->flag); CZK_LAUNCHED();
        CUDA_TRY(ctx, cudaGetLastError());
    }
    CZK_TRY(czk_net_bcast_from_king_dev(ctx, v, bytes));  // "TODO: randomize" in the reference: the king returns the value itself
    return CZK_OK;
}

// open (Net::broadcast + open_degree_vec at every party): out <- value, k elements
static int gsz_open_dev(czk_ctx* ctx, const uint32_t* sh, uint32_t* out, size_t k, int degree) {
    GszState& g = ctx->gsz;
    CZK_TRY(gsz_prepare(ctx));
    g.opens++;
    const size_t bytes = k * 32;
    CZK_TRY(scratch_reserve(ctx, g.gather, bytes * (size_t)ctx->nranks));
    CZK_TRY(czk_net_allgather_dev(ctx, sh, g.gather.p, bytes));
    if (k) {
        k_gsz_open<<<g_grid(k, 256), 256, 0, ctx->stream>>>(out, (const uint32_t*)g.gather.p, k, ctx->nranks, degree, g.winv_dev, g_mk(g.n_inv.l),
                                                            ctx->flag); CZK_LAUNCHED();
        CUDA_TRY(ctx, cudaGetLastError());
    }
    return CZK_OK;
}

// single values: staged through a one-element device buffer so that the scalar steps of the checks take the same path
static int gsz_one_elem(czk_ctx* ctx, uint32_t** p) {
    CZK_TRY(scratch_reserve(ctx, ctx->gsz.one_elem, 64));
    *p = (uint32_t*)ctx->gsz.one_elem.p;
    return CZK_OK;
}
static int gsz_king_compute_host(czk_ctx* ctx, const HFr& v, int degree, HFr* out) {
    uint32_t* d;
    CZK_TRY(gsz_one_elem(ctx, &d));
    CUDA_TRY(ctx, cudaMemcpyAsync(d, v.l, 32, cudaMemcpyHostToDevice, ctx->stream));
    CZK_TRY(gsz_king_compute_dev(ctx, d, 1, degree));
    CUDA_TRY(ctx, cudaMemcpyAsync(out->l, d, 32, cudaMemcpyDeviceToHost, ctx->stream));
    return gsz_check_flag(ctx, "GSZ king_compute");
}
static int gsz_open_host(czk_ctx* ctx, const HFr& v, int degree, HFr* out) {
    uint32_t* d;
    CZK_TRY(gsz_one_elem(ctx, &d));
    CUDA_TRY(ctx, cudaMemcpyAsync(d, v.l, 32, cudaMemcpyHostToDevice, ctx->stream));
    CZK_TRY(gsz_open_dev(ctx, d, d + 8, 1, degree));
    CUDA_TRY(ctx, cudaMemcpyAsync(out->l, d + 8, 32, cudaMemcpyDeviceToHost, ctx->stream));
    return gsz_check_flag(ctx, "GSZ open");
}
int czk_gsz_prepare_internal(czk_ctx* ctx) { return gsz_prepare(ctx); }
int czk_gsz_open_scalar_internal(czk_ctx* ctx, const HFr& v, HFr* out) {
    CZK_TRY(gsz_prepare(ctx));
    return gsz_open_host(ctx, v, ctx->gsz.t, out);
}
// coin (:529-531): open(rand()), rand() = 1 shared at degree t
int gsz_coin(czk_ctx* ctx, HFr* out) {
    CZK_TRY(gsz_prepare(ctx));
    return gsz_open_host(ctx, HFr::one(), ctx->gsz.t, out);
}
// mult (:533-553) of two single shares, not queued (the blinding products of ip_check)
int gsz_mult1(czk_ctx* ctx, const HFr& x, const HFr& y, HFr* out) {
    CZK_TRY(gsz_prepare(ctx));
    HFr v = HFr::add(HFr::mul(x, y), HFr::one());  // + r2
    HFr o;
    CZK_TRY(gsz_king_compute_host(ctx, v, 2 * ctx->gsz.t, &o));
    *out = HFr::sub(o, HFr::one());  // - r
    return CZK_OK;
}

// diagnostics: open_degree_vec on a caller-supplied party-major matrix of gathered shares (parties x k), any party count
// with a share domain - the interpolation and the degree check of k_gsz_open on genuine N-party inputs, on one GPU
int czk_diag_gsz_open_gathered(czk_ctx* ctx, const czk_vec* gathered, int parties, unsigned degree, size_t k, czk_vec* out_pub,
                               uint32_t* flag_out) {
    if (!ctx || !gathered || !out_pub || !flag_out || parties < 1 || parties > 64 || !k || gathered->n < (size_t)parties * k || out_pub->n < k)
        return fail(ctx, CZK_ERR_ARG, "czk_diag_gsz_open_gathered: argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    HFr w;
    if (!gsz_root_of_unity((size_t)parties, &w)) return fail(ctx, CZK_ERR_ARG, "GSZ: no share domain of this size (n must be 2^a or 3 * 2^a)");
    HFr wi = HFr::inv(w), p = HFr::one();
    std::vector<uint64_t> tab((size_t)parties * 4);
    for (int j = 0; j < parties; j++) {
        p.to_limbs(tab.data() + 4 * j);
        p = HFr::mul(p, wi);
    }
    uint32_t* winv = nullptr;
    uint32_t* flag = nullptr;
    CUDA_TRY(ctx, cudaMalloc((void**)&winv, (size_t)parties * 32 + 4));
    flag = winv + (size_t)parties * 8;
    int rc = CZK_OK;
    cudaError_t e = cudaMemcpyAsync(winv, tab.data(), (size_t)parties * 32, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(flag, 0, 4, ctx->stream);
    if (e == cudaSuccess) {
        HFr n_inv = HFr::inv(HFr::from_u64((uint64_t)parties));
        k_gsz_open<<<g_grid(k, 256), 256, 0, ctx->stream>>>((uint32_t*)out_pub->d, (const uint32_t*)gathered->d, k, parties, (int)degree, winv,
                                                            g_mk(n_inv.l), flag); CZK_LAUNCHED();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(flag_out, flag, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, CZK_ERR_CUDA, std::string("czk_diag_gsz_open_gathered: ") + cudaGetErrorString(e));
    cudaFree(winv);
    return rc;
}

// ------------------------------------------------------------------------------------------ C ABI
int czk_gsz_open(czk_ctx* ctx, const czk_vec* sh, unsigned degree, czk_vec* out_pub, size_t n) {
    if (!ctx || !sh || !out_pub || n > sh->n || n > out_pub->n) return fail(ctx, CZK_ERR_ARG, "czk_gsz_open: range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CZK_TRY(gsz_open_dev(ctx, (const uint32_t*)sh->d, (uint32_t*)out_pub->d, n, (int)degree));
    return gsz_check_flag(ctx, "czk_gsz_open");
}

int czk_gsz_king_compute(czk_ctx* ctx, czk_vec* v, unsigned degree, size_t n) {
    if (!ctx || !v || n > v->n) return fail(ctx, CZK_ERR_ARG, "czk_gsz_king_compute: range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CZK_TRY(gsz_king_compute_dev(ctx, (uint32_t*)v->d, n, (int)degree));
    return gsz_check_flag(ctx, "czk_gsz_king_compute");
}

// stream-ordered device buffers that must not outlive an error return
struct AsyncFree {
    cudaStream_t st;
    void* p[3] = {nullptr, nullptr, nullptr};
    ~AsyncFree() {
        for (void* q : p)
            if (q) cudaFreeAsync(q, st);
    }
    void release() { p[0] = p[1] = p[2] = nullptr; }
};

int czk_gsz_batch_mul(czk_ctx* ctx, czk_vec* x, const czk_vec* y, size_t n, int queue_check) {
    if (!ctx || !x || !y || n > x->n || n > y->n) return fail(ctx, CZK_ERR_ARG, "czk_gsz_batch_mul: range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CZK_TRY(gsz_prepare(ctx));
    if (!n) return CZK_OK;
    GszState& g = ctx->gsz;
    GszTriple tr;
    AsyncFree guard{ctx->stream};
    if (queue_check) {  // GszFieldTriple(x, y, z) kept for hadamard_check (:585-592)
        tr.n = n;
        CUDA_TRY(ctx, cudaMallocAsync((void**)&tr.x, n * 32, ctx->stream));
        guard.p[0] = tr.x;
        CUDA_TRY(ctx, cudaMallocAsync((void**)&tr.y, n * 32, ctx->stream));
        guard.p[1] = tr.y;
        CUDA_TRY(ctx, cudaMallocAsync((void**)&tr.z, n * 32, ctx->stream));
        guard.p[2] = tr.z;
        CUDA_TRY(ctx, cudaMemcpyAsync(tr.x, x->d, n * 32, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(tr.y, y->d, n * 32, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    HFr one = HFr::one(), minus_one = HFr::neg(HFr::one());
    // x.val *= y.val; x.val += r2.val   with double_rand() = (1, 1)   (:378-410, :571-577)
    k_gsz_mul_add_const<<<g_grid(n, 256), 256, 0, ctx->stream>>>((uint32_t*)x->d, (const uint32_t*)x->d, (const uint32_t*)y->d, g_mk(one.l), n);
    CZK_LAUNCHED();
    CUDA_TRY(ctx, cudaGetLastError());
    // king just reduces the sharing degree
    CZK_TRY(gsz_king_compute_dev(ctx, (uint32_t*)x->d, n, 2 * g.t));
    CUDA_TRY(ctx, fr_add_const((uint32_t*)x->d, (const uint32_t*)x->d, minus_one.l, n, ctx->stream));  // shift_res.val -= r.val
    if (queue_check) {
        CUDA_TRY(ctx, cudaMemcpyAsync(tr.z, x->d, n * 32, cudaMemcpyDeviceToDevice, ctx->stream));
        g.queue.push_back(tr);
        guard.release();  // the queue owns them now
    }
    return gsz_check_flag(ctx, "czk_gsz_batch_mul");
}

static int gsz_dot(czk_ctx* ctx, const uint32_t* x, const uint32_t* y, size_t h, int mode, const HFr& add, uint32_t* out_dev) {
    GszState& g = ctx->gsz;
    unsigned blocks = g_grid(h, 256);
    CZK_TRY(scratch_reserve(ctx, g.dot_partial, (size_t)blocks * 32));
    k_gsz_dot<<<blocks, 256, 0, ctx->stream>>>((uint32_t*)g.dot_partial.p, x, y, h, mode); CZK_LAUNCHED();
    k_gsz_dot_finish<<<1, 32, 0, ctx->stream>>>(out_dev, (const uint32_t*)g.dot_partial.p, blocks, g_mk(add.l)); CZK_LAUNCHED();
    CUDA_TRY(ctx, cudaGetLastError());
    return CZK_OK;
}
// ip_compute (:738-787): sum_i x_i y_i + r2 -> king_compute at degree 2t -> - r
static int gsz_ip_compute(czk_ctx* ctx, const uint32_t* x, const uint32_t* y, size_t h, int mode, HFr* out) {
    uint32_t* d;
    CZK_TRY(gsz_one_elem(ctx, &d));
    CZK_TRY(gsz_dot(ctx, x, y, h, mode, HFr::one(), d));
    CZK_TRY(gsz_king_compute_dev(ctx, d, 1, 2 * ctx->gsz.t));
    HFr o;
    CUDA_TRY(ctx, cudaMemcpyAsync(o.l, d, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CZK_TRY(gsz_check_flag(ctx, "GSZ ip_compute"));
    *out = HFr::sub(o, HFr::one());
    return CZK_OK;
}

// the parabola through (1, f1), (2, f2), (3, f3) evaluated at r (ip_check, :722-733)
static HFr gsz_parabola(const HFr& v1, const HFr& v2, const HFr& v3, const HFr& r) {
    HFr one = HFr::one(), two = HFr::from_u64(2), three = HFr::from_u64(3), inv2 = HFr::inv(two);
    HFr a = HFr::sub(r, two), b = HFr::sub(r, three), c = HFr::sub(r, one);
    HFr f1 = HFr::mul(HFr::mul(a, b), inv2), f2 = HFr::neg(HFr::mul(c, b)), f3 = HFr::mul(HFr::mul(c, a), inv2);
    return HFr::add(HFr::add(HFr::mul(f1, v1), HFr::mul(f2, v2)), HFr::mul(f3, v3));
}

// hadamard_check (:599-624) -> ip_check (:681-736) over every queued triple; consumes the queue.
// final3 (optional): the opened x | y | z of the last step.
int czk_gsz_check_products(czk_ctx* ctx, uint64_t final3[12]) {
    if (!ctx) return fail(ctx, CZK_ERR_ARG, "czk_gsz_check_products: null");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CZK_TRY(gsz_prepare(ctx));
    GszState& g = ctx->gsz;
    size_t k = 0;
    for (const GszTriple& t : g.queue) k += t.n;
    if (!k) return CZK_OK;
    // working copies with room for the zero pad of odd lengths
    CZK_TRY(scratch_reserve(ctx, g.pad_x, (k + 2) * 32));
    CZK_TRY(scratch_reserve(ctx, g.pad_y, (k + 2) * 32));
    uint32_t *x = (uint32_t*)g.pad_x.p, *y = (uint32_t*)g.pad_y.p;
    uint32_t* z = nullptr;
    CUDA_TRY(ctx, cudaMallocAsync((void**)&z, k * 32, ctx->stream));
    AsyncFree z_guard{ctx->stream};
    z_guard.p[0] = z;
    size_t off = 0;
    for (GszTriple& t : g.queue) {
        CUDA_TRY(ctx, cudaMemcpyAsync(x + off * 8, t.x, t.n * 32, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(y + off * 8, t.y, t.n * 32, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(z + off * 8, t.z, t.n * 32, cudaMemcpyDeviceToDevice, ctx->stream));
        cudaFreeAsync(t.x, ctx->stream);
        cudaFreeAsync(t.y, ctx->stream);
        cudaFreeAsync(t.z, ctx->stream);
        off += t.n;
    }
    g.queue.clear();
    // x_i *= r^i, z_i *= r^i, ip = sum z_i
    HFr r;
    CZK_TRY(gsz_coin(ctx, &r));
    {
        int lo_log = 10;
        size_t nlo = (size_t)1 << lo_log, nhi = (k + nlo - 1) >> lo_log;
        CZK_TRY(scratch_reserve(ctx, ctx->open_sigma, (nlo + nhi) * 32));
        uint32_t* lo = (uint32_t*)ctx->open_sigma.p;
        uint32_t* hi = lo + nlo * 8;
        HFr one = HFr::one();
        CUDA_TRY(ctx, ntt_build_powers(lo, r.l, one.l, nlo, ctx->stream));
        HFr rhi = HFr::pow_u64(r, (uint64_t)nlo);
        CUDA_TRY(ctx, ntt_build_powers(hi, rhi.l, one.l, nhi, ctx->stream));
        CUDA_TRY(ctx, fr_scale_by_tables(x, lo, hi, lo_log, k, ctx->stream));
        CUDA_TRY(ctx, fr_scale_by_tables(z, lo, hi, lo_log, k, ctx->stream));
    }
    uint32_t* d;
    CZK_TRY(gsz_one_elem(ctx, &d));
    CZK_TRY(gsz_dot(ctx, z, nullptr, k, 2, HFr::zero(), d));
    HFr ip;
    CUDA_TRY(ctx, cudaMemcpyAsync(ip.l, d, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    z_guard.release();
    CUDA_TRY(ctx, cudaFreeAsync(z, ctx->stream));
    size_t len = k;
    while (len > 1) {
        if (len & 1) {  // xs.push(from_public(0)); ys.push(from_public(0))
            CUDA_TRY(ctx, cudaMemsetAsync(x + len * 8, 0, 32, ctx->stream));
            CUDA_TRY(ctx, cudaMemsetAsync(y + len * 8, 0, 32, ctx->stream));
            len++;
        }
        const size_t h = len / 2;
        HFr ip_l, ip3, rr;
        CZK_TRY(gsz_ip_compute(ctx, x, y, h, 0, &ip_l));
        HFr ip_r = HFr::sub(ip, ip_l);
        CZK_TRY(gsz_ip_compute(ctx, x, y, h, 1, &ip3));
        CZK_TRY(gsz_coin(ctx, &rr));
        k_gsz_fold<<<g_grid(h, 256), 256, 0, ctx->stream>>>(x, y, h, g_mk(rr.l)); CZK_LAUNCHED();
        CUDA_TRY(ctx, cudaGetLastError());
        ip = gsz_parabola(ip_l, ip_r, ip3, rr);
        len = h;
    }
    // blind with a random pair and open (:722-736): xr = yr = rand() = 1
    HFr x0, y0;
    CUDA_TRY(ctx, cudaMemcpyAsync(x0.l, x, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(y0.l, y, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    HFr one = HFr::one(), ipr, xb, yb, ib, ox, oy, oz;
    CZK_TRY(gsz_mult1(ctx, one, one, &ipr));
    CZK_TRY(gsz_mult1(ctx, x0, one, &xb));
    CZK_TRY(gsz_mult1(ctx, y0, one, &yb));
    CZK_TRY(gsz_mult1(ctx, ip, ipr, &ib));
    CZK_TRY(gsz_open_host(ctx, xb, g.t, &ox));
    CZK_TRY(gsz_open_host(ctx, yb, g.t, &oy));
    CZK_TRY(gsz_open_host(ctx, ib, g.t, &oz));
    ox.to_limbs(g.last_check);
    oy.to_limbs(g.last_check + 4);
    oz.to_limbs(g.last_check + 8);
    if (final3) std::memcpy(final3, g.last_check, sizeof g.last_check);
    if (HFr::mul(ox, oy) != oz) return fail(ctx, CZK_ERR_PROTOCOL, "GSZ product check failed (gsz20/mod.rs:735 assert_eq!(x * y, z))");
    return CZK_OK;
}

int czk_gsz_stats(const czk_ctx* ctx, uint64_t out[2]) {
    if (!ctx || !out) return CZK_ERR_ARG;
    out[0] = ctx->gsz.king_computes;
    out[1] = ctx->gsz.opens;
    return CZK_OK;
}

void gsz_release(czk_ctx* ctx) {
    GszState& g = ctx->gsz;
    for (GszTriple& t : g.queue) {
        cudaFree(t.x);
        cudaFree(t.y);
        cudaFree(t.z);
    }
    g.queue.clear();
    cudaFree(g.winv_dev);
    g.winv_dev = nullptr;
    g.n = 0;
    for (Scratch* s : {&g.dot_partial, &g.one_elem, &g.pad_x, &g.pad_y, &g.gather}) {
        cudaFree(s->p);
        s->p = nullptr;
        s->cap = 0;
    }
}
