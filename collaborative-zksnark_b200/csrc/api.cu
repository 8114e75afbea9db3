// C ABI of libczk_b200.so (include/czk.h): context, device vectors, NTT / MSM orchestration,
// NCCL-backed party network, share opening and Beaver multiplication.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/czk.h"
#include "fr_ops.cuh"
#include "host_field.hpp"
#include "launch_count.hpp"
#include "msm.cuh"
#include "ntt.cuh"

namespace czk {
std::atomic<uint64_t> g_launch_count{0};
cudaError_t microbench_run(int kind, int blocks, int threads, int iters, uint64_t* scratch, double* ops_per_launch,
                           cudaStream_t st);
}  // namespace czk

#include "ctx.hpp"

void gsz_release(czk_ctx* ctx);
const char* czk_version(void) { return "czk-b200 0.1 (sm_100a)"; }

const char* czk_last_error(const czk_ctx* ctx) { return ctx ? ctx->err.c_str() : czk_tls_error().c_str(); }

int czk_ctx_create(int device, czk_ctx** out) {
    if (!out) return fail(nullptr, CZK_ERR_ARG, "czk_ctx_create: out is null");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, CZK_ERR_NO_DEVICE,
                    std::string("no CUDA device available (") + cudaGetErrorString(e) +
                        "); libczk_b200 has no CPU fallback");
    if (device < 0 || device >= count) return fail(nullptr, CZK_ERR_ARG, "czk_ctx_create: device index out of range");
    czk_ctx* ctx = new czk_ctx();
    ctx->device = device;
    CUDA_TRY(ctx, cudaSetDevice(device));
    {
        // the context stream (transforms, share protocols, collectives) outranks the MSM lanes: its short kernels get SM
        // slots as soon as blocks of a long accumulation round retire, instead of queueing behind whole grids
        int lo = 0, hi = 0;
        CUDA_TRY(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        // measured (2^20 SPDZ proof, one B200): 56.5 ms with the context stream prioritised, 55.4 ms without - the witness map
        // finishing early only moves the idle tail; equal priorities stay the default (CZK_STREAM_PRIORITY=1 to try again)
        const char* env = getenv("CZK_STREAM_PRIORITY");
        if (!(env && atoi(env) == 1)) hi = lo;
        CUDA_TRY(ctx, cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, hi));
        ctx->lane_priority = lo;
    }
    {
        cudaMemPool_t pool;
        CUDA_TRY(ctx, cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = ~0ull;
        CUDA_TRY(ctx, cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    CUDA_TRY(ctx, cudaMalloc((void**)&ctx->flag, 4));
    CUDA_TRY(ctx, cudaMemset(ctx->flag, 0, 4));
    for (cudaEvent_t& e : ctx->ev_phase) CUDA_TRY(ctx, cudaEventCreate(&e));
    for (MsmLane& l : ctx->lanes) {
        CUDA_TRY(ctx, cudaStreamCreateWithPriority(&l.stream, cudaStreamNonBlocking, ctx->lane_priority));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&l.ev_in, cudaEventDisableTiming));
        for (MsmSlot& sl : l.slots) {
            CUDA_TRY(ctx, cudaMallocHost(&sl.pinned, 256 * 96 * 4));  // up to 256 XYZZ points of G2
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
            for (cudaEvent_t& e : sl.ev_t) CUDA_TRY(ctx, cudaEventCreate(&e));
        }
    }
    *out = ctx;
    return CZK_OK;
}

static void free_ws(MsmWorkspace& ws) {
    cudaFree(ws.bat_a);
    cudaFree(ws.bat_b);
    cudaFree(ws.bat_prefix);
    cudaFree(ws.scalars);
    cudaFree(ws.hist);
    cudaFree(ws.offsets);
    cudaFree(ws.sorted);
    cudaFree(ws.buckets);
    cudaFree(ws.partial);
    cudaFree(ws.winsum);
    cudaFree(ws.segcnt);
    cudaFree(ws.segoff);
    cudaFree(ws.segsum);
    cudaFree(ws.items);
    cudaFree(ws.heavy);
    cudaFree(ws.queue);
    for (int i = 0; i < 4; i++)
        if (ws.ev[i]) cudaEventDestroy(ws.ev[i]);
    ws = MsmWorkspace();
}

void czk_ctx_destroy(czk_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    czk_net_deinit(ctx);
    for (auto& kv : ctx->domains) {
        Domain& d = kv.second;
        cudaFree(d.tw);
        cudaFree(d.g_lo);
        cudaFree(d.g_hi);
        cudaFree(d.g_hi_sinv);
        cudaFree(d.gi_lo);
        cudaFree(d.gi_hi);
        cudaFree(d.g_sinv_br);
    }
    for (auto& kv : ctx->mixed_domains) {
        cudaFree(kv.second.wpow);
        cudaFree(kv.second.wipow);
    }
    for (MsmLane& l : ctx->lanes) {
        if (l.stream) cudaStreamSynchronize(l.stream);
        for (int i = 0; i < 4; i++) l.ws.ev[i] = nullptr;  // owned by the slots
        free_ws(l.ws);
        for (MsmSlot& sl : l.slots) {
            if (sl.pinned) cudaFreeHost(sl.pinned);
            if (sl.ev_done) cudaEventDestroy(sl.ev_done);
            for (cudaEvent_t e : sl.ev_t)
                if (e) cudaEventDestroy(e);
        }
        if (l.ev_in) cudaEventDestroy(l.ev_in);
        if (l.stream) cudaStreamDestroy(l.stream);
    }
    gsz_release(ctx);
    for (Scratch* s : {&ctx->up_bases, &ctx->up_inf, &ctx->up_scalars, &ctx->up_vec, &ctx->open_gather, &ctx->open_sigma,
                       &ctx->open_sx, &ctx->open_oy, &ctx->open_d, &ctx->open_dm, &ctx->mixed})
        cudaFree(s->p);
    cudaFree(ctx->flag);
    for (cudaEvent_t e : ctx->ev_phase)
        if (e) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int czk_ctx_sync(czk_ctx* ctx) {
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (MsmLane& l : ctx->lanes) CUDA_TRY(ctx, cudaStreamSynchronize(l.stream));
    return CZK_OK;
}
void* czk_ctx_stream(czk_ctx* ctx) { return (void*)ctx->stream; }
uint64_t czk_ctx_launches(const czk_ctx*) { return g_launch_count.load(); }

// ------------------------------------------------------------------------------------------ vectors
int czk_vec_alloc(czk_ctx* ctx, size_t n, czk_vec** out) {
    if (!ctx || !out) return fail(ctx, CZK_ERR_ARG, "czk_vec_alloc: null argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    czk_vec* v = new czk_vec();
    v->n = n;
    size_t bytes = (n ? n : 1) * 32;
    // stream-ordered allocation from the device's default pool (kept, never trimmed: see czk_ctx_create), so the
    // per-proof witness vectors cost no cudaMalloc / cudaFree round trips
    cudaError_t e = cudaMallocAsync((void**)&v->d, bytes, ctx->stream);
    if (e != cudaSuccess) {
        delete v;
        return fail(ctx, CZK_ERR_CUDA, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
    }
    CUDA_TRY(ctx, cudaMemsetAsync(v->d, 0, bytes, ctx->stream));
    *out = v;
    return CZK_OK;
}
void czk_vec_free(czk_ctx* ctx, czk_vec* v) {
    if (!v) return;
    if (ctx) cudaFreeAsync(v->d, ctx->stream);
    else cudaFree(v->d);
    delete v;
}
size_t czk_vec_len(const czk_vec* v) { return v ? v->n : 0; }
uint64_t* czk_vec_device_ptr(czk_vec* v) { return v ? v->d : nullptr; }

int czk_vec_upload(czk_ctx* ctx, czk_vec* v, size_t offset, const uint64_t* host, size_t n) {
    if (!v || (!host && n) || offset + n > v->n) return fail(ctx, CZK_ERR_ARG, "czk_vec_upload: range");
    CUDA_TRY(ctx, cudaMemcpyAsync(v->d + 4 * offset, host, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CZK_OK;
}
int czk_vec_download(czk_ctx* ctx, const czk_vec* v, size_t offset, uint64_t* host, size_t n) {
    if (!v || (!host && n) || offset + n > v->n) return fail(ctx, CZK_ERR_ARG, "czk_vec_download: range");
    CUDA_TRY(ctx, cudaMemcpyAsync(host, v->d + 4 * offset, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CZK_OK;
}
int czk_vec_copy(czk_ctx* ctx, czk_vec* dst, size_t dst_off, const czk_vec* src, size_t src_off, size_t n) {
    if (!dst || !src || dst_off + n > dst->n || src_off + n > src->n) return fail(ctx, CZK_ERR_ARG, "czk_vec_copy: range");
    CUDA_TRY(ctx, cudaMemcpyAsync(dst->d + 4 * dst_off, src->d + 4 * src_off, n * 32, cudaMemcpyDeviceToDevice, ctx->stream));
    return CZK_OK;
}
int czk_vec_zero(czk_ctx* ctx, czk_vec* v, size_t offset, size_t n) {
    if (!v || offset + n > v->n) return fail(ctx, CZK_ERR_ARG, "czk_vec_zero: range");
    CUDA_TRY(ctx, cudaMemsetAsync(v->d + 4 * offset, 0, n * 32, ctx->stream));
    return CZK_OK;
}

// ------------------------------------------------------------------------------------------ NTT
static HFr domain_root(unsigned log_d) {
    HFr w = HFr::from_limbs(FrParams::ROOT_2_47_64);
    for (unsigned i = log_d; i < FrParams::TWO_ADICITY; i++) w = HFr::sqr(w);
    return w;
}

int czk_domain_params(unsigned log_d, uint64_t group_gen[4], uint64_t group_gen_inv[4], uint64_t size_inv[4],
                      uint64_t generator_inv[4]) {
    if (log_d > FrParams::TWO_ADICITY) return fail(nullptr, CZK_ERR_ARG, "domain larger than 2^47");
    HFr w = domain_root(log_d);
    w.to_limbs(group_gen);
    HFr::inv(w).to_limbs(group_gen_inv);
    HFr::inv(HFr::from_u64((uint64_t)1 << log_d)).to_limbs(size_inv);
    HFr::inv(HFr::from_limbs(FrParams::GENERATOR_64)).to_limbs(generator_inv);
    return CZK_OK;
}

static int get_domain(czk_ctx* ctx, int log_d, Domain** out) {
    auto it = ctx->domains.find(log_d);
    if (it != ctx->domains.end()) {
        *out = &it->second;
        return CZK_OK;
    }
    Domain d;
    d.log_d = log_d;
    d.group_gen = domain_root((unsigned)log_d);
    d.group_gen_inv = HFr::inv(d.group_gen);
    d.size_inv = HFr::inv(HFr::from_u64((uint64_t)1 << log_d));
    HFr g = HFr::from_limbs(FrParams::GENERATOR_64);
    d.generator_inv = HFr::inv(g);
    d.lo_log = log_d < 10 ? log_d : 10;
    size_t half = log_d ? ((size_t)1 << (log_d - 1)) : 1;
    size_t nlo = (size_t)1 << d.lo_log, nhi = (size_t)1 << (log_d - d.lo_log);
    CUDA_TRY(ctx, cudaMalloc((void**)&d.tw, (half + 1) * 32));  // omega^k, 0 <= k <= D/2 (the last entry is -1: inverse twiddles read -tw[D/2 - e])
    CUDA_TRY(ctx, cudaMalloc((void**)&d.g_hi_sinv, nhi * 32));
    CUDA_TRY(ctx, cudaMalloc((void**)&d.g_lo, nlo * 32));
    CUDA_TRY(ctx, cudaMalloc((void**)&d.g_hi, nhi * 32));
    CUDA_TRY(ctx, cudaMalloc((void**)&d.gi_lo, nlo * 32));
    CUDA_TRY(ctx, cudaMalloc((void**)&d.gi_hi, nhi * 32));
    HFr one = HFr::one();
    CUDA_TRY(ctx, ntt_build_powers(d.tw, d.group_gen.l, one.l, log_d ? half + 1 : 1, ctx->stream));
    CUDA_TRY(ctx, ntt_build_powers(d.g_lo, g.l, one.l, nlo, ctx->stream));
    HFr g_hi = HFr::pow_u64(g, (uint64_t)nlo);
    CUDA_TRY(ctx, ntt_build_powers(d.g_hi, g_hi.l, one.l, nhi, ctx->stream));
    CUDA_TRY(ctx, ntt_build_powers(d.g_hi_sinv, g_hi.l, d.size_inv.l, nhi, ctx->stream));  // D^-1 g^i: between an iFFT and a coset FFT
    CUDA_TRY(ctx, ntt_build_powers(d.gi_lo, d.generator_inv.l, one.l, nlo, ctx->stream));
    HFr gi_hi = HFr::pow_u64(d.generator_inv, (uint64_t)nlo);
    CUDA_TRY(ctx, ntt_build_powers(d.gi_hi, gi_hi.l, d.size_inv.l, nhi, ctx->stream));
    auto ins = ctx->domains.emplace(log_d, d);
    *out = &ins.first->second;
    return CZK_OK;
}

static NttScale scale_tables(const uint32_t* lo, const uint32_t* hi, int lo_log, bool bitrev) {
    NttScale sc;
    sc.mode = 2;
    sc.bitrev = bitrev ? 1 : 0;
    sc.lo = lo;
    sc.hi = hi;
    sc.lo_log = lo_log;
    return sc;
}
// op: CZK_NTT_FFT / IFFT / COSET_FFT / COSET_IFFT, or CZK_NTT_IFFT_COSET_FFT = ifft_in_place followed by coset_fft_in_place
// (r1cs_to_qap.rs:85-90), which runs as an inverse DIF + a forward DIT with no bit-reversal pass between or after.
static int ntt_batch_dev(czk_ctx* ctx, uint32_t* const* data, int count, unsigned log_d, int op) {
    if (!ctx || (count > 0 && !data) || count < 0) return fail(ctx, CZK_ERR_ARG, "czk_ntt: null argument");
    if (log_d > 30) return fail(ctx, CZK_ERR_ARG, "czk_ntt: log_d > 30 unsupported");
    if (op < 0 || op > CZK_NTT_IFFT_COSET_FFT) return fail(ctx, CZK_ERR_ARG, "czk_ntt: unknown transform");
    if (log_d == 0 || count == 0) return CZK_OK;  // size-1 domain: every transform is the identity
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    Domain* d;
    CZK_TRY(get_domain(ctx, (int)log_d, &d));
    const bool inverse = op == CZK_NTT_IFFT || op == CZK_NTT_COSET_IFFT || op == CZK_NTT_IFFT_COSET_FFT;
    const NttScale none;
    for (int at = 0; at < count; at += NTT_MAX_BATCH) {
        const int cnt = count - at < NTT_MAX_BATCH ? count - at : NTT_MAX_BATCH;
        uint32_t* const* v = data + at;
        if (op == CZK_NTT_IFFT_COSET_FFT && log_d > 2) {
            CUDA_TRY(ctx, ntt_run_tiles(v, cnt, d->tw, (int)log_d, true, false, none, none, ctx->stream));  // natural -> bit-reversed
            // D^-1 g^i on the way into the forward transform; position p holds coefficient bitrev(p).  One table entry per
            // position (32 B x D, built once per domain): a coalesced load and one product where the two-level tables cost two
            // scattered loads and two products
            static const bool direct = [] { const char* e = getenv("CZK_NTT_DIRECT_SCALE"); return !(e && atoi(e) == 0); }();
            if (direct && !d->g_sinv_br && log_d >= 10) {
                HFr g = HFr::from_limbs(FrParams::GENERATOR_64);
                CUDA_TRY(ctx, cudaMalloc((void**)&d->g_sinv_br, ((size_t)1 << log_d) * 32));
                CUDA_TRY(ctx, ntt_build_powers(d->g_sinv_br, g.l, d->size_inv.l, (size_t)1 << log_d, ctx->stream));
                CUDA_TRY(ctx, ntt_bitrev_scale(d->g_sinv_br, (int)log_d, 0, nullptr, nullptr, nullptr, 0, ctx->stream));
            }
            NttScale pre_dit = scale_tables(d->g_lo, d->g_hi_sinv, d->lo_log, true);
            if (d->g_sinv_br) {
                pre_dit = NttScale();
                pre_dit.mode = 3;
                pre_dit.lo = d->g_sinv_br;
            }
            CUDA_TRY(ctx, ntt_run_tiles(v, cnt, d->tw, (int)log_d, false, true, pre_dit, none, ctx->stream));
            continue;
        }
        if (op == CZK_NTT_IFFT_COSET_FFT) {  // <= 4 points: the tiny kernel works in natural order
            NttScale sinv;
            sinv.mode = 1;
            for (int i = 0; i < 4; i++) sinv.c[2 * i] = (uint32_t)d->size_inv.l[i], sinv.c[2 * i + 1] = (uint32_t)(d->size_inv.l[i] >> 32);
            CUDA_TRY(ctx, ntt_run_tiles(v, cnt, d->tw, (int)log_d, true, false, none, sinv, ctx->stream));
            CUDA_TRY(ctx, ntt_run_tiles(v, cnt, d->tw, (int)log_d, false, false, scale_tables(d->g_lo, d->g_hi, d->lo_log, false), none, ctx->stream));
            continue;
        }
        const NttScale pre = op == CZK_NTT_COSET_FFT ? scale_tables(d->g_lo, d->g_hi, d->lo_log, false) : none;
        if (log_d <= 2) {
            NttScale post = none;
            if (op == CZK_NTT_IFFT) {
                post.mode = 1;
                for (int i = 0; i < 4; i++) post.c[2 * i] = (uint32_t)d->size_inv.l[i], post.c[2 * i + 1] = (uint32_t)(d->size_inv.l[i] >> 32);
            } else if (op == CZK_NTT_COSET_IFFT) post = scale_tables(d->gi_lo, d->gi_hi, d->lo_log, false);
            CUDA_TRY(ctx, ntt_run_tiles(v, cnt, d->tw, (int)log_d, inverse, false, pre, post, ctx->stream));
            continue;
        }
        CUDA_TRY(ctx, ntt_run_tiles(v, cnt, d->tw, (int)log_d, inverse, false, pre, none, ctx->stream));
        for (int i = 0; i < cnt; i++) {  // back to natural order; the inverse transforms' scaling rides on the swap
            if (op == CZK_NTT_IFFT) CUDA_TRY(ctx, ntt_bitrev_scale(v[i], (int)log_d, 1, d->size_inv.l, nullptr, nullptr, 0, ctx->stream));
            else if (op == CZK_NTT_COSET_IFFT) CUDA_TRY(ctx, ntt_bitrev_scale(v[i], (int)log_d, 2, nullptr, d->gi_lo, d->gi_hi, d->lo_log, ctx->stream));
            else CUDA_TRY(ctx, ntt_bitrev_scale(v[i], (int)log_d, 0, nullptr, nullptr, nullptr, 0, ctx->stream));
        }
    }
    return CZK_OK;
}

int czk_ntt_fr_dev(czk_ctx* ctx, uint64_t* dev_data, unsigned log_d, int inverse, int coset) {
    if (!ctx || !dev_data) return fail(ctx, CZK_ERR_ARG, "czk_ntt_fr_dev: null argument");
    uint32_t* v = reinterpret_cast<uint32_t*>(dev_data);
    return ntt_batch_dev(ctx, &v, 1, log_d, (inverse ? 1 : 0) | (coset ? 2 : 0));
}

int czk_ntt_fr_batch(czk_ctx* ctx, uint64_t* const* dev_vecs, int count, unsigned log_d, int op) {
    if (!ctx || (count > 0 && !dev_vecs) || count < 0) return fail(ctx, CZK_ERR_ARG, "czk_ntt_fr_batch: null argument");
    std::vector<uint32_t*> v((size_t)count);
    for (int i = 0; i < count; i++) {
        if (!dev_vecs[i]) return fail(ctx, CZK_ERR_ARG, "czk_ntt_fr_batch: null vector");
        v[(size_t)i] = reinterpret_cast<uint32_t*>(dev_vecs[i]);
        for (int j = 0; j < i; j++)
            if (dev_vecs[j] == dev_vecs[i]) return fail(ctx, CZK_ERR_ARG, "czk_ntt_fr_batch: the same vector twice");
    }
    return ntt_batch_dev(ctx, v.data(), count, log_d, op);
}

int czk_ntt_vec_batch(czk_ctx* ctx, czk_vec* const* vecs, int count, unsigned log_d, int op) {
    if (!ctx || (count > 0 && !vecs) || count < 0) return fail(ctx, CZK_ERR_ARG, "czk_ntt_vec_batch: null argument");
    std::vector<uint64_t*> v((size_t)count);
    for (int i = 0; i < count; i++) {
        if (!vecs[i] || vecs[i]->n < ((size_t)1 << log_d)) return fail(ctx, CZK_ERR_ARG, "czk_ntt_vec_batch: vector shorter than the domain");
        v[(size_t)i] = vecs[i]->d;
    }
    return czk_ntt_fr_batch(ctx, v.data(), count, log_d, op);
}

// ------------------------------------------------------------------------------------------ mixed-radix NTT (3 * 2^k)
// FftField::get_root_of_unity(3 * 2^log_m) (algebra/ff/src/fields/mod.rs:337-367, large-subgroup branch)
static HFr mixed_root(unsigned log_m) {
    HFr w = HFr::from_limbs(FrParams::LARGE_ROOT_64);
    for (unsigned i = log_m; i < FrParams::TWO_ADICITY; i++) w = HFr::sqr(w);
    return w;
}
int czk_mixed_domain_params(unsigned log_m, uint64_t group_gen[4], uint64_t group_gen_inv[4], uint64_t size_inv[4],
                            uint64_t generator_inv[4]) {
    if (log_m > FrParams::TWO_ADICITY) return fail(nullptr, CZK_ERR_ARG, "domain larger than 3 * 2^47");
    HFr w = mixed_root(log_m);
    w.to_limbs(group_gen);
    HFr::inv(w).to_limbs(group_gen_inv);
    HFr::inv(HFr::from_u64((uint64_t)3 << log_m)).to_limbs(size_inv);
    HFr::inv(HFr::from_limbs(FrParams::GENERATOR_64)).to_limbs(generator_inv);
    return CZK_OK;
}
static int get_mixed_domain(czk_ctx* ctx, int log_m, MixedDomain** out) {
    auto it = ctx->mixed_domains.find(log_m);
    if (it != ctx->mixed_domains.end()) {
        *out = &it->second;
        return CZK_OK;
    }
    MixedDomain d;
    d.log_m = log_m;
    const size_t M = (size_t)1 << log_m;
    d.group_gen = mixed_root((unsigned)log_m);
    d.group_gen_inv = HFr::inv(d.group_gen);
    d.size_inv = HFr::inv(HFr::from_u64((uint64_t)3 << log_m));
    d.third_inv = HFr::inv(HFr::from_u64(3));
    d.zeta = HFr::pow_u64(d.group_gen, (uint64_t)M);
    d.zeta_inv = HFr::inv(d.zeta);
    HFr one = HFr::one();
    CUDA_TRY(ctx, cudaMalloc((void**)&d.wpow, M * 32));
    CUDA_TRY(ctx, cudaMalloc((void**)&d.wipow, M * 32));
    CUDA_TRY(ctx, ntt_build_powers(d.wpow, d.group_gen.l, one.l, M, ctx->stream));
    CUDA_TRY(ctx, ntt_build_powers(d.wipow, d.group_gen_inv.l, one.l, M, ctx->stream));
    auto ins = ctx->mixed_domains.emplace(log_m, d);
    *out = &ins.first->second;
    return CZK_OK;
}
// a[i] *= c g^i for i < n on a raw device pointer (two-level power table built on the fly)
static int distribute_powers_dev(czk_ctx* ctx, uint32_t* a, const HFr& g, const HFr& c, size_t n) {
    if (!n) return CZK_OK;
    const int lo_log = 10;
    size_t nlo = (size_t)1 << lo_log, nhi = (n + nlo - 1) >> lo_log;
    CZK_TRY(scratch_reserve(ctx, ctx->open_sigma, (nlo + nhi) * 32));
    uint32_t* lo = (uint32_t*)ctx->open_sigma.p;
    uint32_t* hi = lo + nlo * 8;
    HFr one = HFr::one();
    CUDA_TRY(ctx, ntt_build_powers(lo, g.l, one.l, nlo, ctx->stream));
    HFr ghi = HFr::pow_u64(g, (uint64_t)nlo);
    CUDA_TRY(ctx, ntt_build_powers(hi, ghi.l, c.l, nhi, ctx->stream));
    CUDA_TRY(ctx, fr_scale_by_tables(a, lo, hi, lo_log, n, ctx->stream));
    return CZK_OK;
}
// The four transforms of MixedRadixEvaluationDomain (mixed_radix.rs:130-157, domain/mod.rs:139-142) over 3 * 2^log_m
// points for `count` vectors: de-interleave -> 3 * count radix-2 transforms in one batch -> combine (ntt.cu).
// CZK_NTT_IFFT_COSET_FFT runs as the two transforms one after the other.
static int ntt_mixed_batch_dev(czk_ctx* ctx, uint32_t* const* data, int count, unsigned log_m, int op) {
    if (!ctx || (count > 0 && !data) || count < 0) return fail(ctx, CZK_ERR_ARG, "czk_ntt_mixed: null argument");
    if (log_m > 28) return fail(ctx, CZK_ERR_ARG, "czk_ntt_mixed: log_m > 28 unsupported");
    if (op < 0 || op > CZK_NTT_IFFT_COSET_FFT) return fail(ctx, CZK_ERR_ARG, "czk_ntt_mixed: unknown transform");
    if (count == 0) return CZK_OK;
    if (op == CZK_NTT_IFFT_COSET_FFT) {
        CZK_TRY(ntt_mixed_batch_dev(ctx, data, count, log_m, CZK_NTT_IFFT));
        return ntt_mixed_batch_dev(ctx, data, count, log_m, CZK_NTT_COSET_FFT);
    }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    MixedDomain* d;
    CZK_TRY(get_mixed_domain(ctx, (int)log_m, &d));
    const size_t M = (size_t)1 << log_m, N = 3 * M;
    const bool inverse = op == CZK_NTT_IFFT || op == CZK_NTT_COSET_IFFT;
    const HFr g = HFr::from_limbs(FrParams::GENERATOR_64), one = HFr::one();
    CZK_TRY(scratch_reserve(ctx, ctx->mixed, (size_t)count * N * 32));
    uint32_t* sub = (uint32_t*)ctx->mixed.p;
    std::vector<uint32_t*> parts((size_t)count * 3);
    for (int i = 0; i < count; i++) {
        if (op == CZK_NTT_COSET_FFT) CZK_TRY(distribute_powers_dev(ctx, data[i], g, one, N));
        CUDA_TRY(ctx, ntt_mixed_split(data[i], sub + (size_t)i * N * 8, M, ctx->stream));
        for (int r = 0; r < 3; r++) parts[(size_t)i * 3 + r] = sub + ((size_t)i * N + (size_t)r * M) * 8;
    }
    // the radix-2 inverse scales by M^-1; the remaining 3^-1 rides on the combining pass
    CZK_TRY(ntt_batch_dev(ctx, parts.data(), count * 3, log_m, inverse ? CZK_NTT_IFFT : CZK_NTT_FFT));
    for (int i = 0; i < count; i++) {
        CUDA_TRY(ctx, ntt_mixed_combine(sub + (size_t)i * N * 8, data[i], inverse ? d->wipow : d->wpow, inverse ? d->zeta_inv.l : d->zeta.l,
                                        inverse ? d->third_inv.l : nullptr, M, ctx->stream));
        if (op == CZK_NTT_COSET_IFFT) CZK_TRY(distribute_powers_dev(ctx, data[i], HFr::inv(g), one, N));
    }
    return CZK_OK;
}
int czk_ntt_mixed_fr_batch(czk_ctx* ctx, uint64_t* const* dev_vecs, int count, unsigned log_m, int op) {
    if (!ctx || (count > 0 && !dev_vecs) || count < 0) return fail(ctx, CZK_ERR_ARG, "czk_ntt_mixed_fr_batch: null argument");
    std::vector<uint32_t*> v((size_t)count);
    for (int i = 0; i < count; i++) {
        if (!dev_vecs[i]) return fail(ctx, CZK_ERR_ARG, "czk_ntt_mixed_fr_batch: null vector");
        v[(size_t)i] = reinterpret_cast<uint32_t*>(dev_vecs[i]);
        for (int j = 0; j < i; j++)
            if (dev_vecs[j] == dev_vecs[i]) return fail(ctx, CZK_ERR_ARG, "czk_ntt_mixed_fr_batch: the same vector twice");
    }
    return ntt_mixed_batch_dev(ctx, v.data(), count, log_m, op);
}
int czk_ntt_mixed_vec_batch(czk_ctx* ctx, czk_vec* const* vecs, int count, unsigned log_m, int op) {
    if (!ctx || (count > 0 && !vecs) || count < 0) return fail(ctx, CZK_ERR_ARG, "czk_ntt_mixed_vec_batch: null argument");
    if (log_m > 28) return fail(ctx, CZK_ERR_ARG, "czk_ntt_mixed_vec_batch: log_m > 28 unsupported");
    std::vector<uint64_t*> v((size_t)count);
    for (int i = 0; i < count; i++) {
        if (!vecs[i] || vecs[i]->n < ((size_t)3 << log_m)) return fail(ctx, CZK_ERR_ARG, "czk_ntt_mixed_vec_batch: vector shorter than the domain");
        v[(size_t)i] = vecs[i]->d;
    }
    return czk_ntt_mixed_fr_batch(ctx, v.data(), count, log_m, op);
}

// host vector of 3 * 2^log_m plain field elements, in place (what MixedRadixEvaluationDomain::fft_in_place::<Fr> does)
int czk_ntt_mixed_fr(czk_ctx* ctx, uint64_t* host_data, unsigned log_m, int inverse, int coset) {
    if (!ctx || !host_data) return fail(ctx, CZK_ERR_ARG, "czk_ntt_mixed_fr: null argument");
    if (log_m > 28) return fail(ctx, CZK_ERR_ARG, "czk_ntt_mixed_fr: log_m > 28 unsupported");
    size_t bytes = ((size_t)3 << log_m) * 32;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CZK_TRY(scratch_reserve(ctx, ctx->up_vec, bytes));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->up_vec.p, host_data, bytes, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t* v = (uint32_t*)ctx->up_vec.p;
    CZK_TRY(ntt_mixed_batch_dev(ctx, &v, 1, log_m, (inverse ? 1 : 0) | (coset ? 2 : 0)));
    CUDA_TRY(ctx, cudaMemcpyAsync(host_data, ctx->up_vec.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CZK_OK;
}

int czk_ntt_vec(czk_ctx* ctx, czk_vec* v, unsigned log_d, int inverse, int coset) {
    if (!v || v->n < ((size_t)1 << log_d)) return fail(ctx, CZK_ERR_ARG, "czk_ntt_vec: vector shorter than the domain");
    return czk_ntt_fr_dev(ctx, v->d, log_d, inverse, coset);
}

int czk_ntt_fr(czk_ctx* ctx, uint64_t* host_data, unsigned log_d, int inverse, int coset) {
    if (!ctx || !host_data) return fail(ctx, CZK_ERR_ARG, "czk_ntt_fr: null argument");
    if (log_d > 30) return fail(ctx, CZK_ERR_ARG, "czk_ntt_fr: log_d > 30 unsupported");
    size_t bytes = ((size_t)1 << log_d) * 32;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CZK_TRY(scratch_reserve(ctx, ctx->up_vec, bytes));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->up_vec.p, host_data, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CZK_TRY(czk_ntt_fr_dev(ctx, (uint64_t*)ctx->up_vec.p, log_d, inverse, coset));
    CUDA_TRY(ctx, cudaMemcpyAsync(host_data, ctx->up_vec.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CZK_OK;
}

// ------------------------------------------------------------------------------------------ pointwise
#define VEC_CHECK2(a, b, n, name) \
    if (!ctx || !(a) || !(b) || (n) > (a)->n || (n) > (b)->n) return fail(ctx, CZK_ERR_ARG, name ": range")
int czk_vec_add(czk_ctx* ctx, czk_vec* a, const czk_vec* b, size_t n) {
    VEC_CHECK2(a, b, n, "czk_vec_add");
    CUDA_TRY(ctx, fr_binop((uint32_t*)a->d, (const uint32_t*)b->d, n, FR_ADD, ctx->stream));
    return CZK_OK;
}
int czk_vec_sub(czk_ctx* ctx, czk_vec* a, const czk_vec* b, size_t n) {
    VEC_CHECK2(a, b, n, "czk_vec_sub");
    CUDA_TRY(ctx, fr_binop((uint32_t*)a->d, (const uint32_t*)b->d, n, FR_SUB, ctx->stream));
    return CZK_OK;
}
int czk_vec_mul(czk_ctx* ctx, czk_vec* a, const czk_vec* b, size_t n) {
    VEC_CHECK2(a, b, n, "czk_vec_mul");
    CUDA_TRY(ctx, fr_binop((uint32_t*)a->d, (const uint32_t*)b->d, n, FR_MUL, ctx->stream));
    return CZK_OK;
}
int czk_vec_scale(czk_ctx* ctx, czk_vec* a, const uint64_t c[4], size_t n) {
    if (!ctx || !a || !c || n > a->n) return fail(ctx, CZK_ERR_ARG, "czk_vec_scale: range");
    CUDA_TRY(ctx, fr_scale((uint32_t*)a->d, c, n, ctx->stream));
    return CZK_OK;
}
int czk_vec_distribute_powers(czk_ctx* ctx, czk_vec* a, const uint64_t g[4], const uint64_t c[4], size_t n) {
    if (!ctx || !a || !g || !c || n > a->n) return fail(ctx, CZK_ERR_ARG, "czk_vec_distribute_powers: range");
    if (!n) return CZK_OK;
    // a[i] *= c g^i with a two-level power table built on the fly
    int lo_log = 10;
    size_t nlo = (size_t)1 << lo_log, nhi = (n + nlo - 1) >> lo_log;
    CZK_TRY(scratch_reserve(ctx, ctx->open_sigma, (nlo + nhi) * 32));
    uint32_t* lo = (uint32_t*)ctx->open_sigma.p;
    uint32_t* hi = lo + nlo * 8;
    HFr gg = HFr::from_limbs(g), one = HFr::one();
    CUDA_TRY(ctx, ntt_build_powers(lo, gg.l, one.l, nlo, ctx->stream));
    HFr ghi = HFr::pow_u64(gg, (uint64_t)nlo);
    CUDA_TRY(ctx, ntt_build_powers(hi, ghi.l, c, nhi, ctx->stream));
    // reuse the NTT pre-scale kernel on the first n elements (n need not be a power of two)
    CUDA_TRY(ctx, fr_scale_by_tables((uint32_t*)a->d, lo, hi, lo_log, n, ctx->stream));
    return CZK_OK;
}
int czk_vec_divide_by_vanishing_on_coset(czk_ctx* ctx, czk_vec* a, unsigned log_d) {
    size_t n = (size_t)1 << log_d;
    if (!ctx || !a || n > a->n) return fail(ctx, CZK_ERR_ARG, "czk_vec_divide_by_vanishing_on_coset: range");
    // 1 / (g^D - 1)   (radix2/mod.rs:191-193, domain/mod.rs:184-191)
    HFr g = HFr::from_limbs(FrParams::GENERATOR_64);
    HFr z = HFr::sub(HFr::pow_u64(g, (uint64_t)n), HFr::one());
    HFr zi = HFr::inv(z);
    CUDA_TRY(ctx, fr_scale((uint32_t*)a->d, zi.l, n, ctx->stream));
    return CZK_OK;
}

// ------------------------------------------------------------------------------------------ MSM
static int ws_reserve(czk_ctx* ctx, MsmLane& lane, int curve, size_t n, const MsmConfig& cfg) {
    MsmWorkspace& ws = lane.ws;
    cudaStream_t lane_stream = lane.stream;
    size_t total = (size_t)cfg.bwin * cfg.nb;
    int pw = (int)msm_point_words(curve);
    bool grow_n = n > ws.cap_n;
    bool grow_b = total > ws.cap_buckets || pw > ws.point_words;
    if (grow_n) {
        CUDA_TRY(ctx, cudaStreamSynchronize(lane_stream));
        cudaFree(ws.scalars);
        cudaFree(ws.sorted);
        ws.scalars = ws.sorted = nullptr;
        ws.cap_n = 0;  // capacities are recorded only after every allocation of the group succeeded
        ws.cap_sorted_bytes = 0;
        ws.alloc_epoch++;
        size_t cap = n + n / 16 + 64;
        CUDA_TRY(ctx, cudaMalloc((void**)&ws.scalars, cap * 32));
        // worst case windows for this n: ceil(254/2) covers every config
        CUDA_TRY(ctx, cudaMalloc((void**)&ws.sorted, cap * 4 * 32));
        ws.cap_n = cap;
        ws.cap_sorted_bytes = cap * 4 * 32;
    }
    // `sorted` holds n * nwin entries; the allocation above budgets 32 windows, small-c configs need more
    if ((size_t)cfg.nwin > 32) {
        size_t need = n * cfg.nwin * 4;
        if (need > ws.cap_sorted_bytes) {
            CUDA_TRY(ctx, cudaStreamSynchronize(lane_stream));
            cudaFree(ws.sorted);
            ws.sorted = nullptr;
            ws.cap_sorted_bytes = 0;
            ws.alloc_epoch++;
            CUDA_TRY(ctx, cudaMalloc((void**)&ws.sorted, need));
            ws.cap_sorted_bytes = need;
        }
    }
    if (grow_b) {
        CUDA_TRY(ctx, cudaStreamSynchronize(lane_stream));
        cudaFree(ws.hist);
        cudaFree(ws.offsets);
        cudaFree(ws.buckets);
        cudaFree(ws.partial);
        cudaFree(ws.winsum);
        cudaFree(ws.segcnt);
        cudaFree(ws.segoff);
        ws.hist = ws.offsets = ws.buckets = ws.partial = ws.winsum = ws.segcnt = ws.segoff = nullptr;
        ws.alloc_epoch++;
        size_t cap = total > ws.cap_buckets ? total : ws.cap_buckets;
        int pww = pw > ws.point_words ? pw : ws.point_words;
        ws.cap_buckets = 0;
        ws.point_words = 0;
        CUDA_TRY(ctx, cudaMalloc((void**)&ws.hist, cap * 4));
        CUDA_TRY(ctx, cudaMalloc((void**)&ws.offsets, cap * 4));
        CUDA_TRY(ctx, cudaMalloc((void**)&ws.buckets, cap * pww * 4));
        // >= nwin * nchunks points, and room for the bit-slice partial sums of the merged form (<= 24 * (2048 + 32) points)
        CUDA_TRY(ctx, cudaMalloc((void**)&ws.partial, (cap > 65536 ? cap : 65536) * pww * 4));
        CUDA_TRY(ctx, cudaMalloc((void**)&ws.winsum, 256 * (size_t)pww * 4));
        CUDA_TRY(ctx, cudaMalloc((void**)&ws.segcnt, cap * 4));
        CUDA_TRY(ctx, cudaMalloc((void**)&ws.segoff, cap * 4));
        ws.cap_buckets = cap;
        ws.point_words = pww;
    }
    {
        // per-segment sums: one slot per bucket plus one per `seg` sorted entries (see msm_run)
        size_t items = total + (n * cfg.nwin) / 32 + 64;
        if (items > ws.cap_items || pw > ws.seg_point_words) {
            CUDA_TRY(ctx, cudaStreamSynchronize(lane_stream));
            cudaFree(ws.segsum);
            cudaFree(ws.items);
            cudaFree(ws.heavy);
            ws.segsum = ws.items = ws.heavy = nullptr;
            ws.alloc_epoch++;
            size_t cap = items + items / 8;
            int pww = pw > ws.seg_point_words ? pw : ws.seg_point_words;
            ws.cap_items = 0;
            ws.seg_point_words = 0;
            CUDA_TRY(ctx, cudaMalloc((void**)&ws.segsum, cap * (size_t)pww * 4));
            CUDA_TRY(ctx, cudaMalloc((void**)&ws.items, cap * 16));
            CUDA_TRY(ctx, cudaMalloc((void**)&ws.heavy, cap * 4));
            if (!ws.queue) {
                CUDA_TRY(ctx, cudaMalloc((void**)&ws.queue, 64));
                const char* env = getenv("CZK_BATCHED");
                if (!ws.batched_forced) ws.batched = !(env && atoi(env) == 0);
                cudaDeviceProp prop;
                CUDA_TRY(ctx, cudaGetDeviceProperties(&prop, ctx->device));
                ws.sm_count = prop.multiProcessorCount;
            }
            ws.cap_items = cap;
            ws.seg_point_words = pww;
        }
    }
    if (ws.batched) {
        size_t pa, pb, pre;
        msm_batched_bytes(curve, n * cfg.nwin, total, &pa, &pb, &pre);
        struct { uint32_t** p; size_t* cap; size_t need; } bufs[3] = {{&ws.bat_a, &ws.cap_bat_a, pa}, {&ws.bat_b, &ws.cap_bat_b, pb},
                                                                      {&ws.bat_prefix, &ws.cap_bat_prefix, pre}};
        for (auto& b : bufs) {
            if (b.need <= *b.cap) continue;
            CUDA_TRY(ctx, cudaStreamSynchronize(lane_stream));
            cudaFree(*b.p);
            *b.p = nullptr;
            *b.cap = 0;
            size_t want = b.need + b.need / 16;
            CUDA_TRY(ctx, cudaMalloc((void**)b.p, want));
            *b.cap = want;
        }
    }
    return CZK_OK;
}

template <class HF, int LIMBS>
static void msm_host_tail(const uint32_t* winsums, const MsmConfig& cfg, uint64_t* out_xyz) {
    typedef HPoint<HF> P;
    const uint64_t* w64 = reinterpret_cast<const uint64_t*>(winsums);
    auto load = [&](unsigned w) {
        P p;
        const uint64_t* b = w64 + (size_t)w * 4 * LIMBS;
        p.x = HF::from_limbs(b);
        p.y = HF::from_limbs(b + LIMBS);
        p.zz = HF::from_limbs(b + 2 * LIMBS);
        p.zzz = HF::from_limbs(b + 3 * LIMBS);
        return p;
    };
    P total = P::infinity();
    if (msm_uses_bit_sums(cfg)) {
        // winsum = the bit-slice sums S_j of the single (merged) bucket set: sum_j 2^j S_j by Horner
        for (int j = (int)cfg.c - 1; j >= 0; j--) {
            total = P::dbl(total);
            total.add(load((unsigned)j));
        }
    } else {
        for (int w = (int)cfg.bwin - 1; w >= 0; w--) {
            for (unsigned k = 0; k < cfg.c; k++) total = P::dbl(total);
            total.add(load((unsigned)w));
        }
    }
    HF ax, ay;
    if (total.to_affine(ax, ay)) {
        ax.to_limbs(out_xyz);
        ay.to_limbs(out_xyz + LIMBS);
        HF::one().to_limbs(out_xyz + 2 * LIMBS);
    } else {
        // GroupProjective::zero() = (1, 1, 0)
        HF::one().to_limbs(out_xyz);
        HF::one().to_limbs(out_xyz + LIMBS);
        HF::zero().to_limbs(out_xyz + 2 * LIMBS);
    }
}

// Enqueue one MSM on a lane: everything up to the window sums landing in the lane's pinned buffer.  No host
// synchronisation (the number of halving rounds is decided on the device), so the caller can put several MSMs - and other
// work on the context stream - in flight before collecting.  The lane's work is ordered after whatever is already
// enqueued on the context stream (the scalars may come from there).
static int msm_enqueue(czk_ctx* ctx, int lane_idx, int curve, const uint32_t* bases, const uint8_t* inf, const uint32_t* scalars,
                       int mont, size_t n, const MsmConfig* merged_cfg, bool reuse_plan, MsmJob* job) {
    if (n >= ((size_t)1 << 31)) return fail(ctx, CZK_ERR_ARG, "msm: more than 2^31 - 1 terms");
    if (lane_idx < 0 || lane_idx >= CZK_MSM_LANES) return fail(ctx, CZK_ERR_ARG, "msm: lane");
    MsmLane& lane = ctx->lanes[lane_idx];
    if (lane.enqueued - lane.collected >= CZK_MSM_SLOTS) return fail(ctx, CZK_ERR_ARG, "msm: too many uncollected jobs on this lane");
    MsmSlot& slot = lane.slots[lane.enqueued % CZK_MSM_SLOTS];
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    MsmConfig cfg = merged_cfg ? *merged_cfg : msm_choose_config(n ? n : 1);
    const uint64_t epoch = lane.ws.alloc_epoch;
    CZK_TRY(ws_reserve(ctx, lane, curve, n, cfg));
    if (lane.ws.alloc_epoch != epoch) reuse_plan = false;  // a plan buffer moved: the previous plan is gone
    CUDA_TRY(ctx, cudaEventRecord(lane.ev_in, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(lane.stream, lane.ev_in, 0));
    for (int i = 0; i < 4; i++) lane.ws.ev[i] = slot.ev_t[i];
    CUDA_TRY(ctx, msm_run(curve, bases, inf, scalars, mont != 0, n, cfg, lane.ws, lane.stream, reuse_plan));
    size_t pw = msm_point_words(curve);
    CUDA_TRY(ctx, cudaMemcpyAsync(slot.pinned, lane.ws.winsum, msm_winsum_points(cfg) * pw * 4, cudaMemcpyDeviceToHost, lane.stream));
    CUDA_TRY(ctx, cudaEventRecord(slot.ev_done, lane.stream));
    job->seq = lane.enqueued++;
    job->lane = lane_idx;
    job->curve = curve;
    job->n = n;
    job->cfg = cfg;
    return CZK_OK;
}
// Wait for a job's window sums and finish on the host: sum_w 2^(cw) W_w (or the bit-slice Horner) + one affine normalisation.
int msm_collect(czk_ctx* ctx, MsmJob* job, uint64_t* out_xyz, double* device_ms) {
    if (!job || job->lane < 0) return fail(ctx, CZK_ERR_ARG, "msm: no job to collect");
    MsmLane& lane = ctx->lanes[job->lane];
    if (job->seq != lane.collected) return fail(ctx, CZK_ERR_ARG, "msm: jobs of a lane are collected in the order they were enqueued");
    MsmSlot& slot = lane.slots[job->seq % CZK_MSM_SLOTS];
    CUDA_TRY(ctx, cudaEventSynchronize(slot.ev_done));
    lane.collected++;
    {
        float a = 0, m = 0;
        if (cudaEventElapsedTime(&a, slot.ev_t[0], slot.ev_t[1]) == cudaSuccess &&
            cudaEventElapsedTime(&m, slot.ev_t[2], slot.ev_t[3]) == cudaSuccess) {
            int k = job->curve == 1 ? 0 : 1;
            ctx->acc_ms[k] += a;
            ctx->msm_ms[k] += m;
            ctx->acc_terms[k] += (double)job->n;
            ctx->acc_entries[k] += (double)job->n * job->cfg.nwin;
            ctx->acc_launches[k] += 1;
            if (device_ms) *device_ms = m;
        }
    }
    if (job->curve == 1) msm_host_tail<HFq, 6>((const uint32_t*)slot.pinned, job->cfg, out_xyz);
    else msm_host_tail<HFq2, 12>((const uint32_t*)slot.pinned, job->cfg, out_xyz);
    job->lane = -1;
    return CZK_OK;
}
static int msm_core(czk_ctx* ctx, int curve, const uint32_t* bases, const uint8_t* inf, const uint32_t* scalars, int mont,
                    size_t n, uint64_t* out_xyz, const MsmConfig* merged_cfg = nullptr, bool reuse_plan = false) {
    MsmJob job;
    CZK_TRY(msm_enqueue(ctx, 0, curve, bases, inf, scalars, mont, n, merged_cfg, reuse_plan, &job));
    return msm_collect(ctx, &job, out_xyz, nullptr);
}

static int msm_host_entry(czk_ctx* ctx, int curve, const uint64_t* bases_xy, const uint8_t* inf, const uint64_t* scalars,
                          int mont, size_t n, uint64_t* out_xyz) {
    if (!ctx || !out_xyz || (n && (!bases_xy || !scalars))) return fail(ctx, CZK_ERR_ARG, "czk_msm: null argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    size_t pb = curve == 1 ? 96 : 192;
    CZK_TRY(scratch_reserve(ctx, ctx->up_bases, (n ? n : 1) * pb));
    CZK_TRY(scratch_reserve(ctx, ctx->up_scalars, (n ? n : 1) * 32));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->up_bases.p, bases_xy, n * pb, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->up_scalars.p, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    const uint8_t* dinf = nullptr;
    if (inf) {
        CZK_TRY(scratch_reserve(ctx, ctx->up_inf, n ? n : 1));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->up_inf.p, inf, n, cudaMemcpyHostToDevice, ctx->stream));
        dinf = (const uint8_t*)ctx->up_inf.p;
    }
    return msm_core(ctx, curve, (const uint32_t*)ctx->up_bases.p, dinf, (const uint32_t*)ctx->up_scalars.p, mont, n, out_xyz);
}

int czk_msm_g1(czk_ctx* ctx, const uint64_t* bases_xy, const uint8_t* inf, const uint64_t* scalars, int scalars_montgomery,
               size_t n, uint64_t out_xyz[18]) {
    return msm_host_entry(ctx, 1, bases_xy, inf, scalars, scalars_montgomery, n, out_xyz);
}
int czk_msm_g2(czk_ctx* ctx, const uint64_t* bases_xy, const uint8_t* inf, const uint64_t* scalars, int scalars_montgomery,
               size_t n, uint64_t out_xyz[36]) {
    return msm_host_entry(ctx, 2, bases_xy, inf, scalars, scalars_montgomery, n, out_xyz);
}

int czk_bases_upload(czk_ctx* ctx, int curve, const uint64_t* bases_xy, const uint8_t* inf, size_t n, czk_bases** out) {
    if (!ctx || !out || (curve != 1 && curve != 2) || (n && !bases_xy)) return fail(ctx, CZK_ERR_ARG, "czk_bases_upload: argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    czk_bases* b = new czk_bases();
    b->curve = curve;
    b->n = n;
    size_t pb = curve == 1 ? 96 : 192;
    CUDA_TRY(ctx, cudaMalloc((void**)&b->xy, (n ? n : 1) * pb));
    CUDA_TRY(ctx, cudaMemcpyAsync(b->xy, bases_xy, n * pb, cudaMemcpyHostToDevice, ctx->stream));
    bool any = false;
    if (inf)
        for (size_t i = 0; i < n && !any; i++) any = inf[i] != 0;
    if (any) {
        CUDA_TRY(ctx, cudaMalloc((void**)&b->inf, n));
        CUDA_TRY(ctx, cudaMemcpyAsync(b->inf, inf, n, cudaMemcpyHostToDevice, ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *out = b;
    return CZK_OK;
}
void czk_bases_free(czk_ctx* ctx, czk_bases* b) {
    if (!b) return;
    if (ctx) czk_ctx_sync(ctx);  // an MSM lane may still be reading the points
    cudaFree(b->xy);
    cudaFree(b->inf);
    cudaFree(b->table);
    delete b;
}
size_t czk_bases_len(const czk_bases* b) { return b ? b->n : 0; }

int czk_bases_precompute(czk_ctx* ctx, czk_bases* b, unsigned c) {
    if (!ctx || !b) return fail(ctx, CZK_ERR_ARG, "czk_bases_precompute: null argument");
    if (b->n < 1024) return CZK_OK;  // small MSMs keep the windowed form
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (c == 0) c = msm_merged_window(b->n);
    if (c < 4 || c > 24) return fail(ctx, CZK_ERR_ARG, "czk_bases_precompute: window size out of range");
    unsigned nwin = msm_num_windows(c);
    if ((size_t)nwin * b->n >= ((size_t)1 << 31)) return fail(ctx, CZK_ERR_ARG, "czk_bases_precompute: table too large to index");
    if (b->table) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(b->table);
        b->table = nullptr;
    }
    size_t pb = b->curve == 1 ? 96 : 192;
    const unsigned ts = msm_table_stride_words(b->curve);
    CUDA_TRY(ctx, cudaMalloc((void**)&b->table, (size_t)nwin * b->n * ts * 4));
    if (ts * 4 != pb) CUDA_TRY(ctx, cudaMemsetAsync(b->table, 0, (size_t)nwin * b->n * ts * 4, ctx->stream));  // padding words
    // infinity bases must read as (0, 0) so the doubling chain leaves them alone; their scalars are zeroed anyway
    CUDA_TRY(ctx, msm_precompute_table(b->curve, b->table, ts, b->xy, b->n, c, nwin, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    b->pre_c = c;
    b->pre_w = nwin;
    return CZK_OK;
}

int czk_bases_device_bytes(const czk_bases* b, uint64_t out[2]) {
    if (!b || !out) return CZK_ERR_ARG;
    const uint64_t pb = b->curve == 1 ? 96 : 192;
    out[0] = (uint64_t)b->n * pb + (b->inf ? b->n : 0);
    out[1] = b->table ? (uint64_t)b->pre_w * b->n * msm_table_stride_words(b->curve) * 4 : 0;
    return CZK_OK;
}

// czk_msm_bases without the wait: the MSM is enqueued on `lane` (0 or 1) and runs concurrently with the other lane and with
// the context stream; msm_collect returns its result.  Internal (ctx.hpp): the Groth16 prover keeps its five MSMs and the
// witness map in flight together.
int msm_bases_enqueue(czk_ctx* ctx, int lane, const czk_bases* b, size_t base_off, const czk_vec* sc, size_t sc_off,
                      int scalars_montgomery, size_t n, MsmJob* job, bool reuse_plan) {
    if (!ctx || !b || !sc || !job || base_off + n > b->n || sc_off + n > sc->n) return fail(ctx, CZK_ERR_ARG, "czk_msm_bases: range");
    size_t pw = b->curve == 1 ? 24 : 48;
    if (b->table && n >= 1024) {
        MsmConfig cfg = msm_merged_config(b->pre_c, b->n, base_off, msm_table_stride_words(b->curve));
        return msm_enqueue(ctx, lane, b->curve, b->table, b->inf ? b->inf + base_off : nullptr, (const uint32_t*)(sc->d + 4 * sc_off),
                           scalars_montgomery, n, &cfg, reuse_plan, job);
    }
    return msm_enqueue(ctx, lane, b->curve, b->xy + base_off * pw, b->inf ? b->inf + base_off : nullptr,
                       (const uint32_t*)(sc->d + 4 * sc_off), scalars_montgomery, n, nullptr, false, job);
}
int czk_msm_bases(czk_ctx* ctx, const czk_bases* b, size_t base_off, const czk_vec* sc, size_t sc_off, int scalars_montgomery,
                  size_t n, uint64_t* out_xyz) {
    if (!out_xyz) return fail(ctx, CZK_ERR_ARG, "czk_msm_bases: range");
    MsmJob job;
    CZK_TRY(msm_bases_enqueue(ctx, 0, b, base_off, sc, sc_off, scalars_montgomery, n, &job));
    return msm_collect(ctx, &job, out_xyz, nullptr);
}

// Several MSMs over one scalar vector: out[k] = sum_i scalars[sc_off + i] * b[k][base_off + i].  The first base set runs
// in full; a later one reuses its plan (digit decomposition + bucket sort, ~0.5 ms at 2^20 terms) when it addresses a
// precomputed table of the same shape and carries the same infinity flags, and runs as an MSM of its own otherwise.
int czk_msm_bases_multi(czk_ctx* ctx, const czk_bases* const* b, int count, size_t base_off, const czk_vec* sc, size_t sc_off,
                        int scalars_montgomery, size_t n, uint64_t* const* out_xyz, double* ms_out) {
    if (!ctx || !b || !out_xyz || count < 1) return fail(ctx, CZK_ERR_ARG, "czk_msm_bases_multi: argument");
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now();
    auto lap = [&](int k) {
        double t1 = now();
        if (ms_out) ms_out[k] = t1 - t0;
        t0 = t1;
    };
    for (int k = 0; k < count; k++)
        if (!b[k] || !out_xyz[k]) return fail(ctx, CZK_ERR_ARG, "czk_msm_bases_multi: null entry");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const czk_bases* lead = b[0];
    const bool tabled = lead->table && n >= 1024;
    if (tabled) {  // size the workspace for every set first, so that no later reservation moves the plan
        for (int k = 0; k < count; k++) {
            if (!b[k]->table || b[k]->pre_c != lead->pre_c || b[k]->n != lead->n) continue;
            MsmConfig cfg = msm_merged_config(lead->pre_c, lead->n, base_off, msm_table_stride_words(b[k]->curve));
            CZK_TRY(ws_reserve(ctx, ctx->lanes[0], b[k]->curve, n, cfg));
        }
    }
    CZK_TRY(czk_msm_bases(ctx, lead, base_off, sc, sc_off, scalars_montgomery, n, out_xyz[0]));
    lap(0);
    for (int k = 1; k < count; k++) {
        const czk_bases* o = b[k];
        bool share = tabled && o->table && o->pre_c == lead->pre_c && o->n == lead->n && base_off + n <= o->n && sc_off + n <= sc->n &&
                     (o->inf != nullptr) == (lead->inf != nullptr);
        if (share && o->inf) {
            CUDA_TRY(ctx, cudaMemsetAsync(ctx->flag, 0, 4, ctx->stream));
            CUDA_TRY(ctx, msm_flags_differ(o->inf + base_off, lead->inf + base_off, n, ctx->flag, ctx->stream));
            uint32_t differ = 1;
            CUDA_TRY(ctx, cudaMemcpyAsync(&differ, ctx->flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            if (differ) {  // its other users expect it clear
                CUDA_TRY(ctx, cudaMemsetAsync(ctx->flag, 0, 4, ctx->stream));
            }
            share = differ == 0;
        }
        if (!share) {
            // the workspace then holds THAT set's plan, while later sets are compared with the lead's flags: stop sharing
            for (; k < count; k++) {
                CZK_TRY(czk_msm_bases(ctx, b[k], base_off, sc, sc_off, scalars_montgomery, n, out_xyz[k]));
                lap(k);
            }
            return CZK_OK;
        }
        MsmConfig cfg = msm_merged_config(o->pre_c, o->n, base_off, msm_table_stride_words(o->curve));
        CZK_TRY(msm_core(ctx, o->curve, o->table, o->inf ? o->inf + base_off : nullptr, (const uint32_t*)(sc->d + 4 * sc_off),
                         scalars_montgomery, n, out_xyz[k], &cfg, true));
        lap(k);
    }
    return CZK_OK;
}

static uint64_t splitmix(uint64_t& x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int czk_bases_synthetic(czk_ctx* ctx, int curve, uint64_t seed, size_t n, size_t inf_every, czk_bases** out) {
    if (!ctx || !out || (curve != 1 && curve != 2)) return fail(ctx, CZK_ERR_ARG, "czk_bases_synthetic: argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    uint64_t st = seed;
    uint64_t k0[4], ks[4], kq[4], kq2[4];
    for (int i = 0; i < 4; i++) {
        k0[i] = splitmix(st);
        ks[i] = splitmix(st);
    }
    for (int i = 0; i < 4; i++) kq[i] = splitmix(st);
    k0[3] &= (1ull << 56) - 1;  // < 2^248 < r
    ks[3] &= (1ull << 56) - 1;
    kq[3] &= (1ull << 56) - 1;
    ks[0] |= 1;
    kq[0] |= 1;
    for (int i = 0; i < 4; i++) kq2[i] = (kq[i] << 1) | (i ? kq[i - 1] >> 63 : 0);  // 2 kq < 2^249
    czk_bases* b = new czk_bases();
    b->curve = curve;
    b->n = n;
    size_t pb = curve == 1 ? 96 : 192;
    CUDA_TRY(ctx, cudaMalloc((void**)&b->xy, (n ? n : 1) * pb + 2 * pb));
    // generator and second-difference point (2 kquad * G) on the host, staged after the output array
    std::vector<uint64_t> gs(2 * pb / 8);
    if (curve == 1) {
        HG1 g = HG1::from_affine(HFq::from_limbs(CurveConsts::G1_GEN), HFq::from_limbs(CurveConsts::G1_GEN + 6));
        HG1 s = HG1::mul(g, kq2, 4);
        HFq sx, sy;
        s.to_affine(sx, sy);
        std::memcpy(gs.data(), CurveConsts::G1_GEN, 96);
        sx.to_limbs(gs.data() + 12);
        sy.to_limbs(gs.data() + 18);
    } else {
        HG2 g = HG2::from_affine(HFq2::from_limbs(CurveConsts::G2_GEN), HFq2::from_limbs(CurveConsts::G2_GEN + 12));
        HG2 s = HG2::mul(g, kq2, 4);
        HFq2 sx, sy;
        s.to_affine(sx, sy);
        std::memcpy(gs.data(), CurveConsts::G2_GEN, 192);
        sx.to_limbs(gs.data() + 24);
        sy.to_limbs(gs.data() + 36);
    }
    uint32_t* stage = b->xy + (n ? n : 1) * (pb / 4);
    CUDA_TRY(ctx, cudaMemcpyAsync(stage, gs.data(), 2 * pb, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, ec_gen_progression_dev(curve, b->xy, stage, stage + pb / 4, k0, ks, kq, n, ctx->stream));
    if (inf_every && n) {
        std::vector<uint8_t> flags(n, 0);
        for (size_t i = inf_every - 1; i < n; i += inf_every) flags[i] = 1;
        CUDA_TRY(ctx, cudaMalloc((void**)&b->inf, n));
        CUDA_TRY(ctx, cudaMemcpyAsync(b->inf, flags.data(), n, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *out = b;
    return CZK_OK;
}

int czk_bases_download(czk_ctx* ctx, const czk_bases* b, size_t off, size_t n, uint64_t* xy, uint8_t* inf) {
    if (!ctx || !b || off + n > b->n || (n && !xy)) return fail(ctx, CZK_ERR_ARG, "czk_bases_download: range");
    size_t pb = b->curve == 1 ? 96 : 192;
    CUDA_TRY(ctx, cudaMemcpyAsync(xy, (const uint8_t*)b->xy + off * pb, n * pb, cudaMemcpyDeviceToHost, ctx->stream));
    if (inf) {
        if (b->inf) CUDA_TRY(ctx, cudaMemcpyAsync(inf, b->inf + off, n, cudaMemcpyDeviceToHost, ctx->stream));
        else std::memset(inf, 0, n);
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CZK_OK;
}

void gsz_release(czk_ctx* ctx);
// ------------------------------------------------------------------------------------------ network
int czk_net_unique_id(uint8_t out[128]) {
    NcclApi& api = nccl_api();
    if (!api.ok) return fail(nullptr, CZK_ERR_NCCL, "libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    ncclResult_t r = api.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, CZK_ERR_NCCL, std::string("ncclGetUniqueId: ") + api.GetErrorString(r));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    std::memcpy(out, &id, 128);
    return CZK_OK;
}
int czk_net_init(czk_ctx* ctx, int rank, int nranks, const uint8_t nccl_unique_id[128]) {
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, CZK_ERR_ARG, "czk_net_init: rank");
    czk_net_deinit(ctx);
    ctx->rank = rank;
    ctx->nranks = nranks;
    if (nranks == 1) return CZK_OK;
    NcclApi& api = nccl_api();
    if (!api.ok) return fail(ctx, CZK_ERR_NCCL, "libnccl.so.2 could not be loaded");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    std::memcpy(&id, nccl_unique_id, 128);
    ncclResult_t r = api.CommInitRank(&ctx->comm, nranks, id, rank);
    if (r != ncclSuccess) return fail(ctx, CZK_ERR_NCCL, std::string("ncclCommInitRank: ") + api.GetErrorString(r));
    return CZK_OK;
}
int czk_net_init_single(czk_ctx* ctx) { return czk_net_init(ctx, 0, 1, nullptr); }
void czk_net_deinit(czk_ctx* ctx) {
    if (ctx) sh_p2p_release(ctx);
    if (ctx && ctx->comm) {
        cudaStreamSynchronize(ctx->stream);
        nccl_api().CommDestroy(ctx->comm);
        ctx->comm = nullptr;
    }
    if (ctx) {
        ctx->rank = 0;
        ctx->nranks = 1;
    }
}
int czk_net_party_id(const czk_ctx* ctx) { return ctx->rank; }
int czk_net_n_parties(const czk_ctx* ctx) { return ctx->nranks; }

int czk_net_allgather_dev(czk_ctx* ctx, const void* dev_send, void* dev_recv, size_t bytes) {
    if (!ctx || !dev_send || !dev_recv) return fail(ctx, CZK_ERR_ARG, "czk_net_allgather_dev: null");
    // mpc-net broadcast accounting (multi.rs:145-174): one message to, and one from, every other party
    ctx->stats[0] += bytes * (uint64_t)(ctx->nranks - 1);
    ctx->stats[1] += bytes * (uint64_t)(ctx->nranks - 1);
    ctx->stats[2] += 1;
    if (ctx->nranks == 1) {
        if (dev_send != dev_recv) CUDA_TRY(ctx, cudaMemcpyAsync(dev_recv, dev_send, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        return CZK_OK;
    }
    ncclResult_t r = nccl_api().AllGather(dev_send, dev_recv, bytes, ncclUint8, ctx->comm, ctx->stream);
    if (r != ncclSuccess) return fail(ctx, CZK_ERR_NCCL, std::string("ncclAllGather: ") + nccl_api().GetErrorString(r));
    return CZK_OK;
}
int czk_net_allgather_host(czk_ctx* ctx, const void* host_send, void* host_recv, size_t bytes) {
    if (!ctx || !host_send || !host_recv) return fail(ctx, CZK_ERR_ARG, "czk_net_allgather_host: null");
    size_t total = bytes * (size_t)ctx->nranks;
    CZK_TRY(scratch_reserve(ctx, ctx->open_d, bytes + total + 32));
    uint8_t* send = (uint8_t*)ctx->open_d.p;
    uint8_t* recv = send + ((bytes + 15) / 16) * 16;
    CUDA_TRY(ctx, cudaMemcpyAsync(send, host_send, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CZK_TRY(czk_net_allgather_dev(ctx, send, recv, bytes));
    CUDA_TRY(ctx, cudaMemcpyAsync(host_recv, recv, total, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CZK_OK;
}
int czk_net_bcast_from_king_dev(czk_ctx* ctx, void* dev_buf, size_t bytes) {
    if (!ctx || !dev_buf) return fail(ctx, CZK_ERR_ARG, "czk_net_bcast_from_king_dev: null");
    if (ctx->rank == 0) ctx->stats[0] += bytes * (uint64_t)(ctx->nranks - 1);
    else ctx->stats[1] += bytes;
    ctx->stats[4] += 1;
    if (ctx->nranks == 1) return CZK_OK;
    ncclResult_t r = nccl_api().Broadcast(dev_buf, dev_buf, bytes, ncclUint8, 0, ctx->comm, ctx->stream);
    if (r != ncclSuccess) return fail(ctx, CZK_ERR_NCCL, std::string("ncclBroadcast: ") + nccl_api().GetErrorString(r));
    return CZK_OK;
}
int czk_net_stats(const czk_ctx* ctx, uint64_t out[5]) {
    for (int i = 0; i < 5; i++) out[i] = ctx->stats[i];
    return CZK_OK;
}
void czk_net_reset_stats(czk_ctx* ctx) {
    for (int i = 0; i < 5; i++) ctx->stats[i] = 0;
}

// shares: czk_batch_open / czk_beaver_batch_mul live in shares.cu

// ------------------------------------------------------------------------------------------ diagnostics
int czk_msm_set_batched(czk_ctx* ctx, int enabled) {
    if (!ctx) return CZK_ERR_ARG;
    CZK_TRY(czk_ctx_sync(ctx));
    for (MsmLane& l : ctx->lanes) {
        l.ws.batched = enabled != 0;
        l.ws.batched_always = enabled == 2;
        l.ws.batched_forced = true;
    }
    return CZK_OK;
}
int czk_fq_inverse(czk_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n) {
    if (!ctx || !in || !out) return fail(ctx, CZK_ERR_ARG, "czk_fq_inverse: null");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CZK_TRY(scratch_reserve(ctx, ctx->up_bases, n * 96));
    uint32_t* d = (uint32_t*)ctx->up_bases.p;
    CUDA_TRY(ctx, cudaMemcpyAsync(d, in, n * 48, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, fq_inverse_batch(d, d + n * 12, n, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(out, d + n * 12, n * 48, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return CZK_OK;
}
int czk_msm_stats(czk_ctx* ctx, int curve, double out[5], int reset) {
    if (!ctx || !out || (curve != 1 && curve != 2)) return fail(ctx, CZK_ERR_ARG, "czk_msm_stats: argument");
    int k = curve - 1;
    out[0] = ctx->acc_ms[k];
    out[1] = (double)ctx->acc_launches[k];
    out[2] = ctx->acc_terms[k];
    out[3] = ctx->msm_ms[k];
    out[4] = ctx->acc_entries[k];
    if (reset) {
        ctx->acc_ms[k] = ctx->msm_ms[k] = ctx->acc_terms[k] = ctx->acc_entries[k] = 0;
        ctx->acc_launches[k] = 0;
    }
    return CZK_OK;
}

int czk_microbench(czk_ctx* ctx, int kind, int blocks_per_sm, int threads, int iters, double* ops_per_s, double* ms) {
    if (!ctx || !ops_per_s || !ms) return fail(ctx, CZK_ERR_ARG, "czk_microbench: null");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CUDA_TRY(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    int blocks = prop.multiProcessorCount * blocks_per_sm;
    CZK_TRY(scratch_reserve(ctx, ctx->up_vec, (size_t)blocks * threads * 8));
    double ops = 0;
    cudaEvent_t e0, e1;
    CUDA_TRY(ctx, cudaEventCreate(&e0));
    CUDA_TRY(ctx, cudaEventCreate(&e1));
    CUDA_TRY(ctx, microbench_run(kind, blocks, threads, iters, (uint64_t*)ctx->up_vec.p, &ops, ctx->stream));  // warm-up
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
        CUDA_TRY(ctx, microbench_run(kind, blocks, threads, iters, (uint64_t*)ctx->up_vec.p, &ops, ctx->stream));
        CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
        CUDA_TRY(ctx, cudaEventSynchronize(e1));
        float t;
        CUDA_TRY(ctx, cudaEventElapsedTime(&t, e0, e1));
        if (t < best) best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms = best;
    *ops_per_s = ops / (best * 1e-3);
    return CZK_OK;
}

