// Polynomial and share-protocol leaves of the Plonk / KZG10 provers, on device vectors.
//
//   czk_vec_prefix_products     Field::partial_products_in_place on plain values (ff/src/fields/mod.rs:222-234)
//   czk_vec_batch_inverse       batch inversion of public values (the `.inverse().unwrap()` map of share/field.rs:142)
//   czk_poly_div_linear         DensePolynomial::divide_with_q_and_r by (X - z) (poly/src/polynomial/univariate/mod.rs:133-174),
//                               the witness polynomial of KZG10::open (poly-commit/src/kzg10/mod.rs:196-220); on shares it is
//                               applied to every share vector (share/add.rs:148-156)
//   czk_share_batch_inv / _div / _partial_products   FieldShare::{batch_inv, batch_div, partial_products}
//                               (mpc-algebra/src/share/field.rs:135-182) with the stub inv pairs of wire/field.rs:62-77
//   czk_kzg_open                KZG10::open without hiding (kzg10/mod.rs:226-262): evaluation + MSM of the witness coefficients
//
// Both recurrences on the path - prefix products and the Horner chain of a division by (X - z) - are scans.  Prefix products
// are a scan under multiplication.  The division is q_i = sum_{j>i} p_j z^(j-i-1): scale by powers (u_j = p_j z^j), suffix-SUM
// scan, scale back by z^-i - two streaming passes and an additive scan instead of a serial chain of n multiply-adds.
// Every value is a canonical field element, so the results are bit-identical to the reference's serial loops.
#include "ctx.hpp"
#include "fr_ops.cuh"
#include "launch_count.hpp"
#include "ntt.cuh"

namespace czk {

__device__ __forceinline__ Fr p_ld(const uint32_t* p, size_t i) {
    const uint4* q = reinterpret_cast<const uint4*>(p) + 2 * i;
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void p_st(uint32_t* p, size_t i, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p) + 2 * i;
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// ------------------------------------------------------------------ scans (inclusive; MUL: product, else sum)
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
template <bool MUL>
__device__ __forceinline__ Fr scan_op(const Fr& a, const Fr& b) {
    return MUL ? Fr::mul(a, b) : Fr::add(a, b);
}
template <bool MUL>
__device__ __forceinline__ Fr scan_identity() {
    return MUL ? Fr::one() : Fr::zero();
}
// One tile of SCAN_TILE elements per block.  REV: the scan runs from the last element to the first (element i of the
// scanned sequence is data[n - 1 - i]).  aggr[block] = the tile total (may be null for a single tile).
template <bool MUL, bool REV>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                              uint32_t* __restrict__ aggr, size_t n) {
    __shared__ uint32_t sm[2][SCAN_THREADS * 8];
    const unsigned tid = threadIdx.x;
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)tid * SCAN_ITEMS;
    Fr v[SCAN_ITEMS];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        size_t i = base + k;
        v[k] = i < n ? p_ld(in, REV ? n - 1 - i : i) : scan_identity<MUL>();
        if (k) v[k] = scan_op<MUL>(v[k - 1], v[k]);
    }
    // inclusive scan of the thread totals across the block (Hillis-Steele, double buffered, limb-major)
    int cur = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) sm[0][w * SCAN_THREADS + tid] = v[SCAN_ITEMS - 1].l[w];
    __syncthreads();
    for (unsigned off = 1; off < SCAN_THREADS; off <<= 1) {
        Fr a;
#pragma unroll
        for (int w = 0; w < 8; w++) a.l[w] = sm[cur][w * SCAN_THREADS + tid];
        if (tid >= off) {
            Fr b;
#pragma unroll
            for (int w = 0; w < 8; w++) b.l[w] = sm[cur][w * SCAN_THREADS + tid - off];
            a = scan_op<MUL>(b, a);
        }
#pragma unroll
        for (int w = 0; w < 8; w++) sm[cur ^ 1][w * SCAN_THREADS + tid] = a.l[w];
        cur ^= 1;
        __syncthreads();
    }
    Fr excl = scan_identity<MUL>();
    if (tid) {
#pragma unroll
        for (int w = 0; w < 8; w++) excl.l[w] = sm[cur][w * SCAN_THREADS + tid - 1];
    }
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        size_t i = base + k;
        if (i < n) p_st(out, REV ? n - 1 - i : i, tid ? scan_op<MUL>(excl, v[k]) : v[k]);
    }
    if (aggr && tid == SCAN_THREADS - 1) {
        Fr t;
#pragma unroll
        for (int w = 0; w < 8; w++) t.l[w] = sm[cur][w * SCAN_THREADS + tid];
        p_st(aggr, blockIdx.x, t);
    }
}
// out[i] = op(aggr_incl[tile - 1], out[i]) for every tile but the first
template <bool MUL, bool REV>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(uint32_t* __restrict__ out, const uint32_t* __restrict__ aggr_incl, size_t n) {
    if (blockIdx.x == 0) return;
    const Fr pre = p_ld(aggr_incl, blockIdx.x - 1);
    const size_t base = (size_t)blockIdx.x * SCAN_TILE;
    for (int k = 0; k < SCAN_ITEMS; k++) {
        size_t i = base + (size_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) {
            size_t j = REV ? n - 1 - i : i;
            p_st(out, j, scan_op<MUL>(pre, p_ld(out, j)));
        }
    }
}
template <bool MUL, bool REV>
static cudaError_t fr_scan_t(const uint32_t* in, uint32_t* out, size_t n, uint32_t* scratch, cudaStream_t st) {
    if (!n) return cudaSuccess;
    size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_tiles<MUL, REV><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, out, tiles > 1 ? scratch : nullptr, n); CZK_LAUNCHED();
    if (tiles > 1) {
        // scan the tile totals in place (forward), then fold them into the tiles
        cudaError_t e = fr_scan_t<MUL, false>(scratch, scratch, tiles, scratch + tiles * 8, st);
        if (e != cudaSuccess) return e;
        k_scan_apply<MUL, REV><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(out, scratch, n); CZK_LAUNCHED();
    }
    return cudaGetLastError();
}
// scratch: room for the tile totals of every level: < n / 1024 elements
cudaError_t fr_scan(const uint32_t* in, uint32_t* out, size_t n, bool mul, bool reverse, uint32_t* scratch, cudaStream_t st) {
    if (mul) return reverse ? fr_scan_t<true, true>(in, out, n, scratch, st) : fr_scan_t<true, false>(in, out, n, scratch, st);
    return reverse ? fr_scan_t<false, true>(in, out, n, scratch, st) : fr_scan_t<false, false>(in, out, n, scratch, st);
}

// ------------------------------------------------------------------ batch inversion (Montgomery's trick per thread)
constexpr int INV_BATCH = 16;
__global__ void __launch_bounds__(128) k_fr_batch_inverse(uint32_t* __restrict__ a, size_t n, uint32_t* __restrict__ flag) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    // element j of this thread: index t + j * stride (coalesced across the warp)
    Fr pre[INV_BATCH];
    Fr acc = Fr::one();
    bool zero = false;
#pragma unroll
    for (int j = 0; j < INV_BATCH; j++) {
        size_t i = t + (size_t)j * stride;
        Fr x = i < n ? p_ld(a, i) : Fr::one();
        if (x.is_zero()) {
            zero = true;
            x = Fr::one();
        }
        pre[j] = acc;  // product of the elements before j
        acc = Fr::mul(acc, x);
    }
    // acc^(r-2)
    Fr inv = Fr::inv_fermat(acc);
#pragma unroll
    for (int j = INV_BATCH - 1; j >= 0; j--) {
        size_t i = t + (size_t)j * stride;
        if (i < n) {
            Fr x = p_ld(a, i);
            if (x.is_zero()) x = Fr::one();
            p_st(a, i, Fr::mul(inv, pre[j]));
            inv = Fr::mul(inv, x);
        }
    }
    if (__any_sync(0xffffffffu, zero) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}
cudaError_t fr_batch_inverse(uint32_t* a, size_t n, uint32_t* flag, cudaStream_t st) {
    if (!n) return cudaSuccess;
    size_t threads = (n + INV_BATCH - 1) / INV_BATCH;
    k_fr_batch_inverse<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(a, n, flag); CZK_LAUNCHED();
    return cudaGetLastError();
}

// q[i] = s[i + 1] for i < n - 1 (the quotient of the division by X - z), s = the back-scaled suffix sums
__global__ void k_poly_shift_down(uint32_t* __restrict__ q, const uint32_t* __restrict__ s, size_t m) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x) p_st(q, i, p_ld(s, i + 1));
}
// v[i] = c   (constant share vectors of the stub sources)
__global__ void k_fr_fill(uint32_t* __restrict__ v, FrConst c, size_t n) {
    Fr cc;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        cc.l[2 * i] = (uint32_t)c.v[i];
        cc.l[2 * i + 1] = (uint32_t)(c.v[i] >> 32);
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p_st(v, i, cc);
}

}  // namespace czk

using namespace czk;

static unsigned p_grid(size_t n) {
    size_t b = (n + 255) / 256;
    return (unsigned)(b < 148 * 8 ? (b ? b : 1) : 148 * 8);
}
static int p_check_flag(czk_ctx* ctx, const char* what) {
    uint32_t flag = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&flag, ctx->flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (flag) {
        cudaMemsetAsync(ctx->flag, 0, 4, ctx->stream);
        return fail(ctx, CZK_ERR_PROTOCOL, std::string(what) + ": inverse of zero (the reference's .inverse().unwrap() panics)");
    }
    return CZK_OK;
}
static int p_fill(czk_ctx* ctx, czk_vec* v, const HFr& c, size_t n) {
    FrConst fc;
    for (int i = 0; i < 4; i++) fc.v[i] = c.l[i];
    k_fr_fill<<<p_grid(n), 256, 0, ctx->stream>>>((uint32_t*)v->d, fc, n); CZK_LAUNCHED();
    CUDA_TRY(ctx, cudaGetLastError());
    return CZK_OK;
}
// scratch for the scans' tile totals
static int p_scan_scratch(czk_ctx* ctx, size_t n, uint32_t** out) {
    CZK_TRY(scratch_reserve(ctx, ctx->open_sigma, (n / 1024 + 64) * 32));
    *out = (uint32_t*)ctx->open_sigma.p;
    return CZK_OK;
}
// a[i] *= c g^i on a raw device pointer (the two-level power tables of czk_vec_distribute_powers)
static int p_scale_powers(czk_ctx* ctx, uint32_t* a, const HFr& g, size_t n) {
    int lo_log = 10;
    size_t nlo = (size_t)1 << lo_log, nhi = (n + nlo - 1) >> lo_log;
    CZK_TRY(scratch_reserve(ctx, ctx->open_oy, (nlo + nhi) * 32));
    uint32_t* lo = (uint32_t*)ctx->open_oy.p;
    uint32_t* hi = lo + nlo * 8;
    HFr one = HFr::one();
    CUDA_TRY(ctx, ntt_build_powers(lo, g.l, one.l, nlo, ctx->stream));
    HFr ghi = HFr::pow_u64(g, (uint64_t)nlo);
    CUDA_TRY(ctx, ntt_build_powers(hi, ghi.l, one.l, nhi, ctx->stream));
    CUDA_TRY(ctx, fr_scale_by_tables(a, lo, hi, lo_log, n, ctx->stream));
    return CZK_OK;
}

// ------------------------------------------------------------------------------------------ plain-value leaves
int czk_vec_prefix_products(czk_ctx* ctx, czk_vec* v, size_t n) {
    if (!ctx || !v || n > v->n) return fail(ctx, CZK_ERR_ARG, "czk_vec_prefix_products: range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    uint32_t* scr;
    CZK_TRY(p_scan_scratch(ctx, n, &scr));
    CUDA_TRY(ctx, fr_scan((const uint32_t*)v->d, (uint32_t*)v->d, n, true, false, scr, ctx->stream));
    return CZK_OK;
}
int czk_vec_batch_inverse(czk_ctx* ctx, czk_vec* v, size_t n) {
    if (!ctx || !v || n > v->n) return fail(ctx, CZK_ERR_ARG, "czk_vec_batch_inverse: range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, fr_batch_inverse((uint32_t*)v->d, n, ctx->flag, ctx->stream));
    return p_check_flag(ctx, "czk_vec_batch_inverse");
}
// p / (X - z): q (n - 1 coefficients, may alias nothing of p) and rem = p(z) (device, first element of rem_dev if given)
static int poly_div_linear_dev(czk_ctx* ctx, const uint32_t* p, size_t n, const HFr& z, uint32_t* q, uint64_t rem_host[4]) {
    if (!n) {
        if (rem_host) std::memset(rem_host, 0, 32);
        return CZK_OK;
    }
    uint32_t* s = nullptr;  // the Horner values s_i = sum_{j >= i} p_j z^(j-i)
    CUDA_TRY(ctx, cudaMallocAsync((void**)&s, n * 32, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(s, p, n * 32, cudaMemcpyDeviceToDevice, ctx->stream));
    if (!z.is_zero()) {
        uint32_t* scr;
        CZK_TRY(p_scan_scratch(ctx, n, &scr));
        CZK_TRY(p_scale_powers(ctx, s, z, n));                                               // u_j = p_j z^j
        CUDA_TRY(ctx, fr_scan(s, s, n, false, true, scr, ctx->stream));                      // U_i = sum_{j >= i} u_j
        CZK_TRY(p_scale_powers(ctx, s, HFr::inv(z), n));                                     // s_i = U_i z^-i
    }
    // z = 0: s_i = p_i already (q_i = p_{i+1}, rem = p_0)
    if (q && n > 1) {
        k_poly_shift_down<<<p_grid(n - 1), 256, 0, ctx->stream>>>(q, s, n - 1); CZK_LAUNCHED();
        CUDA_TRY(ctx, cudaGetLastError());
    }
    if (rem_host) CUDA_TRY(ctx, cudaMemcpyAsync(rem_host, s, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaFreeAsync(s, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CZK_OK;
}
int czk_poly_div_linear(czk_ctx* ctx, const czk_vec* p, size_t n, const uint64_t z[4], czk_vec* q_out, uint64_t rem_out[4]) {
    if (!ctx || !p || !z || n > p->n || (q_out && n > 1 && q_out->n < n - 1)) return fail(ctx, CZK_ERR_ARG, "czk_poly_div_linear: range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return poly_div_linear_dev(ctx, (const uint32_t*)p->d, n, HFr::from_limbs(z), q_out ? (uint32_t*)q_out->d : nullptr, rem_out);
}

// ------------------------------------------------------------------------------------------ share protocols
struct TmpVec {
    czk_ctx* ctx;
    czk_vec* v = nullptr;
    explicit TmpVec(czk_ctx* c) : ctx(c) {}
    ~TmpVec() { czk_vec_free(ctx, v); }
    int alloc(size_t n) { return czk_vec_alloc(ctx, n, &v); }
};
static HFr king_one(const czk_ctx* ctx) { return ctx->rank == 0 ? HFr::one() : HFr::zero(); }

int czk_share_batch_inv(czk_ctx* ctx, int scheme, czk_vec* x_sh, czk_vec* x_mac, size_t n) {
    if (!ctx || !x_sh || n > x_sh->n) return fail(ctx, CZK_ERR_ARG, "czk_share_batch_inv: range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (scheme == CZK_SCHEME_PLAIN) return czk_vec_batch_inverse(ctx, x_sh, n);
    if (scheme != CZK_SCHEME_ADDITIVE && scheme != CZK_SCHEME_SPDZ) return fail(ctx, CZK_ERR_ARG, "czk_share_batch_inv: scheme");
    const bool spdz = scheme == CZK_SCHEME_SPDZ;
    if (spdz && (!x_mac || n > x_mac->n)) return fail(ctx, CZK_ERR_ARG, "SPDZ inverse needs the MAC vector");
    if (!n) return CZK_OK;
    // (b, c) = inv_pairs: from_add_shared(1 at the king) for both (wire/field.rs:62-77); SPDZ mac = share (key 1)
    TmpVec b(ctx), o(ctx);
    CZK_TRY(b.alloc(n));
    CZK_TRY(o.alloc(n));
    CZK_TRY(p_fill(ctx, b.v, king_one(ctx), n));
    CZK_TRY(czk_beaver_batch_mul(ctx, scheme, x_sh, x_mac, b.v, b.v, n));                    // batch_mul(xs, bs)
    CZK_TRY(czk_batch_open(ctx, scheme, x_sh, x_mac, o.v, n));                                // batch_open
    CUDA_TRY(ctx, fr_batch_inverse((uint32_t*)o.v->d, n, ctx->flag, ctx->stream));            // .inverse().unwrap()
    CZK_TRY(p_check_flag(ctx, "czk_share_batch_inv"));
    // c.scale(&i): c = 1 at the king, 0 elsewhere (value and MAC share alike)
    if (ctx->rank == 0) {
        CZK_TRY(czk_vec_copy(ctx, x_sh, 0, o.v, 0, n));
        if (spdz) CZK_TRY(czk_vec_copy(ctx, x_mac, 0, o.v, 0, n));
    } else {
        CZK_TRY(czk_vec_zero(ctx, x_sh, 0, n));
        if (spdz) CZK_TRY(czk_vec_zero(ctx, x_mac, 0, n));
    }
    return CZK_OK;
}

int czk_share_batch_div(czk_ctx* ctx, int scheme, czk_vec* x_sh, czk_vec* x_mac, czk_vec* y_sh, czk_vec* y_mac, size_t n) {
    if (!ctx || !x_sh || !y_sh || n > x_sh->n || n > y_sh->n) return fail(ctx, CZK_ERR_ARG, "czk_share_batch_div: range");
    // batch_mul(xs, batch_inv(ys))   (share/field.rs:155-158); y returns holding the shares of 1 / y
    CZK_TRY(czk_share_batch_inv(ctx, scheme, y_sh, y_mac, n));
    return czk_beaver_batch_mul(ctx, scheme, x_sh, x_mac, y_sh, y_mac, n);
}

int czk_share_partial_products(czk_ctx* ctx, int scheme, czk_vec* x_sh, czk_vec* x_mac, size_t n) {
    if (!ctx || !x_sh || n > x_sh->n) return fail(ctx, CZK_ERR_ARG, "czk_share_partial_products: range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (scheme == CZK_SCHEME_PLAIN) return czk_vec_prefix_products(ctx, x_sh, n);
    if (scheme != CZK_SCHEME_ADDITIVE && scheme != CZK_SCHEME_SPDZ) return fail(ctx, CZK_ERR_ARG, "czk_share_partial_products: scheme");
    const bool spdz = scheme == CZK_SCHEME_SPDZ;
    if (spdz && (!x_mac || n > x_mac->n)) return fail(ctx, CZK_ERR_ARG, "SPDZ partial products need the MAC vector");
    if (!n) return CZK_OK;
    // share/field.rs:160-182 with every (m_i, m_i^-1) = (1 at the king, 1 at the king)
    TmpVec m(ctx), mm(ctx), c1(ctx), pub(ctx);
    CZK_TRY(m.alloc(n));
    CZK_TRY(mm.alloc(n));
    CZK_TRY(c1.alloc(n));
    CZK_TRY(pub.alloc(n));
    const HFr k1 = king_one(ctx);
    CZK_TRY(p_fill(ctx, m.v, k1, n));
    CZK_TRY(p_fill(ctx, mm.v, k1, n));
    CZK_TRY(p_fill(ctx, c1.v, k1, n));
    czk_vec* mmac = spdz ? mm.v : nullptr;
    CZK_TRY(czk_beaver_batch_mul(ctx, scheme, m.v, mmac, x_sh, x_mac, n));                  // mx = batch_mul(m[..n], x)
    CZK_TRY(czk_beaver_batch_mul(ctx, scheme, m.v, mmac, c1.v, c1.v, n));                   // mxm = batch_mul(mx, m_inv[1..])
    CZK_TRY(czk_batch_open(ctx, scheme, m.v, mmac, pub.v, n));                              // mxm_pub
    CZK_TRY(czk_vec_prefix_products(ctx, pub.v, n));                                        // mxm_pub[i] *= mxm_pub[i-1]
    CZK_TRY(p_fill(ctx, m.v, k1, n));                                                       // m0 = [m[0]; n]
    CZK_TRY(p_fill(ctx, mm.v, k1, n));
    CZK_TRY(czk_beaver_batch_mul(ctx, scheme, m.v, mmac, c1.v, c1.v, n));                   // mms = batch_mul(m0, m_inv[1..])
    CZK_TRY(czk_share_batch_inv(ctx, scheme, m.v, mmac, n));                                // mms_inv
    CUDA_TRY(ctx, fr_binop((uint32_t*)m.v->d, (const uint32_t*)pub.v->d, n, FR_MUL, ctx->stream));   // .scale(&mxm_pub[i])
    CZK_TRY(czk_vec_copy(ctx, x_sh, 0, m.v, 0, n));
    if (spdz) {
        CUDA_TRY(ctx, fr_binop((uint32_t*)mm.v->d, (const uint32_t*)pub.v->d, n, FR_MUL, ctx->stream));
        CZK_TRY(czk_vec_copy(ctx, x_mac, 0, mm.v, 0, n));
    }
    return CZK_OK;
}

// ------------------------------------------------------------------------------------------ KZG10
// open without hiding on one coefficient vector (a plain polynomial, or this party's value / MAC share vector):
// eval = p(z), w = MSM(powers_of_g, coefficients of (p - p(z)) / (X - z))
int czk_kzg_open(czk_ctx* ctx, const czk_bases* powers, const czk_vec* p, size_t n, const uint64_t z[4], uint64_t w_xyz[18],
                 uint64_t eval_out[4]) {
    if (!ctx || !powers || !p || !z || !w_xyz || !eval_out || n > p->n || (n > 1 && n - 1 > czk_bases_len(powers)))
        return fail(ctx, CZK_ERR_ARG, "czk_kzg_open: range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TmpVec q(ctx);
    CZK_TRY(q.alloc(n > 1 ? n - 1 : 1));
    CZK_TRY(poly_div_linear_dev(ctx, (const uint32_t*)p->d, n, HFr::from_limbs(z), (uint32_t*)q.v->d, eval_out));
    return czk_msm_bases(ctx, powers, 0, q.v, 0, 1, n > 1 ? n - 1 : 0, w_xyz);
}
