// Additive / SPDZ share opening and Beaver multiplication: fused kernels around a reduce-scatter-shaped exchange.
//
// Replaces FieldShare::{batch_open, batch_mul} for AdditiveFieldShare / SpdzFieldShare:
//   mpc-algebra/src/share/add.rs:121-125     batch_open  = Net::broadcast + sum
//   mpc-algebra/src/share/spdz.rs:166-185    batch_open  = broadcast + sum, sigma = mac_share * x - mac, atomic_broadcast, assert sum == 0
//   mpc-algebra/src/share/field.rs:97-127    batch_mul   = open(s + x), open(o + y), z - sx*y - oy*x + shift(sx*oy)
//   mpc-algebra/src/wire/field.rs:41-77      DummyFieldTripleSource: x = y = z = from_add_shared(1 at the king, 0 elsewhere)
//
// The reference opens by all-to-all broadcast: every party receives N full vectors and sums them.  Here one open of T
// elements between N parties is
//   pack      own shares (+ this party's triple share) into a send buffer of N slices of m = ceil(T / N) elements,
//   exchange  slice q of every party's buffer goes to party q           (grouped ncclSend / ncclRecv, (N-1)/N * T elements),
//   reduce    party q sums the N copies of slice q                      (k_sh_reduce_slices),
//   gather    the summed slices are all-gathered                        (in-place ncclAllGather, (N-1)/N * T elements),
// so a party moves 2 T elements per open instead of N T, and nobody adds more than T values.  Beaver's two opens are
// independent (d = s + x, e = o + y), so they travel as ONE open of 2n elements; SPDZ's sigma check is a second
// exchange + reduce with nothing to gather (each party checks its slice, the verdicts are all-gathered as one word).  A
// SPDZ product is therefore 4 kernels (pack, reduce, after_open = sigma + Beaver finish for both components, check)
// around 3 collectives and ONE flag read-back; the additive product is 3 kernels around 2 collectives and no
// host synchronisation at all.
//
// The kernels take LISTS of source / destination pointers, one per party.  Over NCCL the list addresses the receive
// buffer; in the single-GPU N-party simulation used by the parity tests (czk_diag_sim_*) it addresses the other
// simulated parties' buffers directly, so every kernel runs on genuine N-party inputs without a second GPU.
#include "ctx.hpp"
#include "fr_ops.cuh"
#include "launch_count.hpp"

namespace czk {

constexpr int SH_MAXP = 16;  // parties of an additive / SPDZ computation (one NVSwitch domain holds 8 GPUs)
struct ShPtrs {
    const uint32_t* p[SH_MAXP];
};
struct ShDsts {
    uint32_t* p[SH_MAXP];
};

__device__ __forceinline__ Fr sh_ld(const uint32_t* p, size_t i) {
    const uint4* q = reinterpret_cast<const uint4*>(p) + 2 * i;
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void sh_st(uint32_t* p, size_t i, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p) + 2 * i;
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ Fr sh_cst(const FrConst& c) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        r.l[2 * i] = (uint32_t)c.v[i];
        r.l[2 * i + 1] = (uint32_t)(c.v[i] >> 32);
    }
    return r;
}
static FrConst sh_mk(const HFr& h) {
    FrConst r;
    for (int i = 0; i < 4; i++) r.v[i] = h.l[i];
    return r;
}
static unsigned sh_grid(size_t n) {
    size_t b = (n + 255) / 256, cap = 148 * 8;  // grid-stride over a few waves of the 148 SMs
    return (unsigned)(b < cap ? (b ? b : 1) : cap);
}
#define SH_STRIDE(i, n) for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (size_t)gridDim.x * blockDim.x)

// send[i] = a[i] + t (i < n), send[n + i] = b[i] + t (b != nullptr), zero padding up to `padded`
__global__ void __launch_bounds__(256) k_sh_pack(uint32_t* __restrict__ send, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                                 FrConst t, size_t n, size_t padded) {
    const Fr tt = sh_cst(t);
    const size_t cnt = b ? 2 * n : n;
    SH_STRIDE(i, padded) {
        Fr v = Fr::zero();
        if (i < n) v = Fr::add(sh_ld(a, i), tt);
        else if (i < cnt) v = Fr::add(sh_ld(b, i - n), tt);
        sh_st(send, i, v);
    }
}

// dst_k[i] = sum_p src_p[i], i < m: the reduce of the reduce-scatter, written to `ndst` destinations (1 over NCCL, where the
// all-gather follows; every party's buffer in the single-GPU simulation)
__global__ void __launch_bounds__(256) k_sh_reduce_slices(ShDsts dst, int ndst, ShPtrs src, int parties, size_t m) {
    SH_STRIDE(i, m) {
        Fr acc = sh_ld(src.p[0], i);
        for (int p = 1; p < parties; p++) acc = Fr::add(acc, sh_ld(src.p[p], i));
        for (int k = 0; k < ndst; k++) sh_st(dst.p[k], i, acc);
    }
}

// What follows an open, fused into one pass over the opened values:
//   SPDZ:    sigma[i] = mac_share * opened[i] - (mac_in[i] + t_mac)                       (spdz.rs:173-178), zero padded
//   product: out[i] = tz - ty * sx[i] - tx * oy[i] + shift * sx[i] * oy[i]  per component  (field.rs:116-126), sx = opened[i],
//            oy = opened[n + i]
// mac_a / mac_b: the MAC vectors of the two opened halves (b unused for a plain open).  out_sh / out_mac may alias the
// inputs: every element is read before it is written by the same thread.
// A per-party protocol constant (triple share, MAC key share).  With the reference's stub preprocessing every one of them
// is 0 or 1, and a product by 0 or 1 needs no multiplication: kind 0 = zero, 1 = one, 2 = any value (general product).
struct ShConst {
    FrConst v;
    int kind;
};
__device__ __forceinline__ Fr sh_mulc(const ShConst& c, const Fr& x) {
    if (c.kind == 0) return Fr::zero();
    if (c.kind == 1) return x;
    return Fr::mul(sh_cst(c.v), x);
}
struct ShTriple {
    ShConst x, y, z, shift;
};
static ShConst sh_mkc(const HFr& h) {
    ShConst c;
    c.v = sh_mk(h);
    c.kind = h.is_zero() ? 0 : (h == HFr::one() ? 1 : 2);
    return c;
}
__global__ void __launch_bounds__(256) k_sh_after_open(const uint32_t* __restrict__ opened, size_t n, size_t padded, int spdz, int product,
                                                       const uint32_t* mac_a, const uint32_t* mac_b, FrConst t_mac, ShConst mac_share,
                                                       uint32_t* __restrict__ sigma, uint32_t* out_sh, uint32_t* out_mac, ShTriple tv,
                                                       ShTriple tm) {
    const Fr tmac = sh_cst(t_mac);
    const ShConst ms = mac_share;
    const size_t cnt = product ? 2 * n : n;
    SH_STRIDE(i, n) {
        const Fr sx = sh_ld(opened, i);
        Fr oy;
        if (product) oy = sh_ld(opened, n + i);
        if (spdz) {
            sh_st(sigma, i, Fr::sub(sh_mulc(ms, sx), Fr::add(sh_ld(mac_a, i), tmac)));
            if (product) sh_st(sigma, n + i, Fr::sub(sh_mulc(ms, oy), Fr::add(sh_ld(mac_b, i), tmac)));
        }
        if (product) {
            const Fr de = Fr::mul(sx, oy);
            // z.sub(y.scale(&sx)).sub(x.scale(&oy)).shift(&(sx * oy))
            Fr r = Fr::sub(sh_cst(tv.z.v), sh_mulc(tv.y, sx));
            r = Fr::sub(r, sh_mulc(tv.x, oy));
            r = Fr::add(r, sh_mulc(tv.shift, de));
            sh_st(out_sh, i, r);
            if (spdz) {
                Fr q = Fr::sub(sh_cst(tm.z.v), sh_mulc(tm.y, sx));
                q = Fr::sub(q, sh_mulc(tm.x, oy));
                q = Fr::add(q, sh_mulc(tm.shift, de));
                sh_st(out_mac, i, q);
            }
        }
    }
    if (spdz) {
        const Fr z = Fr::zero();
        SH_STRIDE(j, padded - cnt) sh_st(sigma, cnt + j, z);
    }
}

// *flag |= 1 if sum_p src_p[i] != 0 for some i < m   (spdz.rs:179-182: assert!(sum.is_zero()))
__global__ void __launch_bounds__(256) k_sh_check_zero(ShPtrs src, int parties, size_t m, uint32_t* __restrict__ flag) {
    uint32_t bad = 0;
    SH_STRIDE(i, m) {
        Fr acc = sh_ld(src.p[0], i);
        for (int p = 1; p < parties; p++) acc = Fr::add(acc, sh_ld(src.p[p], i));
        if (!acc.is_zero()) bad = 1;
    }
    if (bad) atomicOr(flag, 1u);
}

}  // namespace czk

using namespace czk;

// ------------------------------------------------------------------------------------------ one party's view of a job
struct ShJob {
    int rank = 0;
    const uint32_t *a_sh = nullptr, *a_mac = nullptr;  // first opened vector (Beaver: s), n elements
    const uint32_t *b_sh = nullptr, *b_mac = nullptr;  // second opened vector (Beaver: o); nullptr for a plain open
    uint32_t *out_sh = nullptr, *out_mac = nullptr;    // product: result components (may alias a_*); open: out_sh = opened values
    uint32_t *send = nullptr, *recv = nullptr, *opened = nullptr, *sigma = nullptr;  // N * m elements each
    uint32_t* flag = nullptr;                          // device word, this party's check verdict
};
struct ShShape {
    int scheme, parties;
    bool product;
    size_t n, total, m, padded;  // total = n or 2n opened elements; m = slice; padded = parties * m
};
static ShShape sh_shape(int scheme, int parties, bool product, size_t n) {
    ShShape s;
    s.scheme = scheme;
    s.parties = parties;
    s.product = product;
    s.n = n;
    s.total = product ? 2 * n : n;
    s.m = (s.total + (size_t)parties - 1) / (size_t)parties;
    if (s.m == 0) s.m = 1;
    s.padded = s.m * (size_t)parties;
    return s;
}

// mpc-net's own accounting of what the reference would have sent for this open (multi.rs:145-174: a broadcast of b bytes
// counts (N-1) b sent, (N-1) b received, 1 broadcast; Vec<F> serialises as an 8-byte length + 32 bytes per element;
// atomic_broadcast = a 32-byte commitment broadcast + the data with 32 bytes of commitment randomness, channel.rs:50-75),
// so that czk_net_stats() stays comparable with the reference's Stats line.  What actually crossed NVLink is in link_bytes.
static void sh_count_reference_open(czk_ctx* ctx, int scheme, size_t k) {
    const uint64_t peers = (uint64_t)(ctx->nranks - 1);
    uint64_t bytes = 8 + 32 * (uint64_t)k, casts = 1;
    if (scheme == CZK_SCHEME_SPDZ) {
        bytes += 32 + (8 + 32 * (uint64_t)k + 32);
        casts = 3;
    }
    ctx->stats[0] += bytes * peers;
    ctx->stats[1] += bytes * peers;
    ctx->stats[2] += casts;
}

static HFr sh_king_one(int rank) { return rank == 0 ? HFr::one() : HFr::zero(); }

static int sh_launch_pack(czk_ctx* ctx, const ShShape& s, const ShJob& j) {
    // Beaver: the opened values are s + x and o + y with the stub triple share x = y = (1 at the king); a plain open adds 0
    HFr t = s.product ? sh_king_one(j.rank) : HFr::zero();
    k_sh_pack<<<sh_grid(s.padded), 256, 0, ctx->stream>>>(j.send, j.a_sh, j.b_sh, sh_mk(t), s.n, s.padded); CZK_LAUNCHED();
    CUDA_TRY(ctx, cudaGetLastError());
    return CZK_OK;
}
static int sh_launch_after_open(czk_ctx* ctx, const ShShape& s, const ShJob& j) {
    const bool spdz = s.scheme == CZK_SCHEME_SPDZ;
    if (!spdz && !s.product) return CZK_OK;
    const HFr king = sh_king_one(j.rank);
    // stub triple: value shares (1 at the king, 0 elsewhere); SPDZ from_add_shared(f) sets mac = f * mac() = f (spdz.rs:138-143);
    // shift(public) adds at the king for the value share and public * mac_share for the MAC share (mac_share = 1 at the king)
    ShTriple tv{sh_mkc(king), sh_mkc(king), sh_mkc(king), sh_mkc(king)};
    ShTriple tm = tv;
    HFr tmac = s.product ? king : HFr::zero();
    k_sh_after_open<<<sh_grid(s.n), 256, 0, ctx->stream>>>(j.opened, s.n, s.padded, spdz ? 1 : 0, s.product ? 1 : 0, j.a_mac, j.b_mac, sh_mk(tmac),
                                                         sh_mkc(king), j.sigma, j.out_sh, j.out_mac, tv, tm); CZK_LAUNCHED();
    CUDA_TRY(ctx, cudaGetLastError());
    return CZK_OK;
}

// ------------------------------------------------------------------------------------------ over NCCL: this rank is one party
static int sh_nccl_exchange(czk_ctx* ctx, const uint32_t* send, uint32_t* recv, size_t slice_bytes) {
    NcclApi& api = nccl_api();
    ncclResult_t r = api.GroupStart();
    for (int p = 0; p < ctx->nranks && r == ncclSuccess; p++) {
        if (p == ctx->rank) continue;  // the own slice is read in place by the reduce kernel
        r = api.Send((const uint8_t*)send + (size_t)p * slice_bytes, slice_bytes, ncclUint8, p, ctx->comm, ctx->stream);
        if (r == ncclSuccess) r = api.Recv((uint8_t*)recv + (size_t)p * slice_bytes, slice_bytes, ncclUint8, p, ctx->comm, ctx->stream);
    }
    ncclResult_t r2 = api.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) return fail(ctx, CZK_ERR_NCCL, std::string("slice exchange: ") + api.GetErrorString(r));
    ctx->link_bytes[0] += slice_bytes * (uint64_t)(ctx->nranks - 1);
    ctx->link_bytes[1] += slice_bytes * (uint64_t)(ctx->nranks - 1);
    return CZK_OK;
}
static ShPtrs sh_nccl_sources(const czk_ctx* ctx, const uint32_t* send, const uint32_t* recv, size_t m) {
    // every source is the slice addressed to THIS rank: party p's copy was received into slot p of `recv`, the own copy is
    // read where it was packed (slot rank of `send`)
    ShPtrs s{};
    for (int p = 0; p < ctx->nranks; p++) s.p[p] = (p == ctx->rank ? send : recv) + (size_t)p * m * 8;
    return s;
}

// The whole job on this rank.  Leaves the opened values in j.opened[0 .. total) (on every rank) and, for SPDZ, the verdict of
// this rank's slice in *j.flag; the caller reads the flags back (sh_collect_flags) when it needs the answer.
static int sh_reserve(czk_ctx* ctx, const ShShape& s, ShJob& j);
static int sh_run_nccl(czk_ctx* ctx, const ShShape& s, ShJob& j) {
    const int N = ctx->nranks;
    CZK_TRY(sh_reserve(ctx, s, j));
    const bool spdz = s.scheme == CZK_SCHEME_SPDZ;
    const size_t slice_bytes = s.m * 32;
    CZK_TRY(sh_launch_pack(ctx, s, j));
    if (N == 1) {
        j.opened = j.send;  // the sum over one party
    } else {
        CZK_TRY(sh_nccl_exchange(ctx, j.send, j.recv, slice_bytes));
        ShDsts d{};
        d.p[0] = j.opened + (size_t)ctx->rank * s.m * 8;
        k_sh_reduce_slices<<<sh_grid(s.m), 256, 0, ctx->stream>>>(d, 1, sh_nccl_sources(ctx, j.send, j.recv, s.m), N, s.m); CZK_LAUNCHED();
        CUDA_TRY(ctx, cudaGetLastError());
        ncclResult_t r = nccl_api().AllGather((const uint8_t*)j.opened + (size_t)ctx->rank * slice_bytes, j.opened, slice_bytes, ncclUint8,
                                              ctx->comm, ctx->stream);
        if (r != ncclSuccess) return fail(ctx, CZK_ERR_NCCL, std::string("ncclAllGather: ") + nccl_api().GetErrorString(r));
        ctx->link_bytes[0] += slice_bytes * (uint64_t)(N - 1);
        ctx->link_bytes[1] += slice_bytes * (uint64_t)(N - 1);
    }
    CZK_TRY(sh_launch_after_open(ctx, s, j));
    if (spdz) {
        ShPtrs src{};
        if (N == 1) {
            src.p[0] = j.sigma;
        } else {
            CZK_TRY(sh_nccl_exchange(ctx, j.sigma, j.recv, slice_bytes));
            src = sh_nccl_sources(ctx, j.sigma, j.recv, s.m);
        }
        k_sh_check_zero<<<sh_grid(s.m), 256, 0, ctx->stream>>>(src, N, s.m, j.flag); CZK_LAUNCHED();
        CUDA_TRY(ctx, cudaGetLastError());
    }
    return CZK_OK;
}

// ------------------------------------------------------------------------------------------ over NVLink peer memory
// Inside one NVSwitch domain every rank can address every other rank's HBM.  The exchange buffers are then allocated once,
// exported with CUDA IPC and mapped by every peer; an open needs NO bulk collective: the reduce kernel of rank q reads slice
// q of every peer's send buffer through its NVLink mapping and writes the sum into every peer's `opened` buffer - the
// reduce-scatter and the all-gather of the NCCL path fused into one kernel - and the sigma check reads the peers' sigma
// slices the same way.  What remains of NCCL is a 4-byte all-gather between the kernels, used as the barrier that orders
// "every rank has packed" before "any rank reads", and "every rank has written my opened buffer" before I read it.
//   pack -> barrier -> reduce (P2P loads, P2P stores) -> barrier -> after_open -> barrier -> check (P2P loads)
// Buffer reuse across consecutive opens is safe by the same barriers (a rank joins barrier 1 of open k + 1 only after its
// own kernels of open k, so nobody's send / sigma buffer is rewritten while a peer still reads it).
// CZK_SHARE_TRANSPORT=nccl keeps the send/recv exchange; the probe falls back to it when any rank cannot map any peer.
static int sh_allgather_host(czk_ctx* ctx, const void* send, void* recv, size_t bytes) { return czk_net_allgather_host(ctx, send, recv, bytes); }

void sh_p2p_release(czk_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    bool mapped = false;
    for (P2PBuf* b : {&ctx->p2p_send, &ctx->p2p_opened, &ctx->p2p_sigma})
        for (int q = 0; q < CZK_P2P_MAX; q++)
            if (b->peer[q] && b->peer[q] != b->local) {
                cudaIpcCloseMemHandle(b->peer[q]);
                b->peer[q] = nullptr;
                mapped = true;
            }
    if (mapped && ctx->comm && ctx->nranks > 1) {  // every rank has unmapped before anybody frees (deinit is collective)
        int token = 1, tokens[CZK_P2P_MAX];
        czk_net_allgather_host(ctx, &token, tokens, sizeof(int));
    }
    for (P2PBuf* b : {&ctx->p2p_send, &ctx->p2p_opened, &ctx->p2p_sigma}) {
        cudaFree(b->local);
        *b = P2PBuf();
    }
    cudaFree(ctx->p2p_sync);
    ctx->p2p_sync = nullptr;
    ctx->p2p_state = 0;
}

// decide once per communicator, collectively: every rank must be able to reach every other rank's device
static int sh_p2p_probe(czk_ctx* ctx) {
    if (ctx->p2p_state != 0) return CZK_OK;
    const int N = ctx->nranks;
    int ok = N > 1 && N <= CZK_P2P_MAX;
    const char* env = getenv("CZK_SHARE_TRANSPORT");
    if (env && std::string(env) == "nccl") ok = 0;
    int devs[CZK_P2P_MAX] = {0};
    int mine = ctx->device;
    if (N > 1) CZK_TRY(sh_allgather_host(ctx, &mine, devs, sizeof(int)));
    for (int q = 0; q < N && ok; q++) {
        if (q == ctx->rank) continue;
        if (devs[q] == ctx->device) {  // two ranks on one device cannot use IPC mappings of each other
            ok = 0;
            break;
        }
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, ctx->device, devs[q]) != cudaSuccess || !can) ok = 0;
    }
    int all[CZK_P2P_MAX] = {0};
    if (N > 1) CZK_TRY(sh_allgather_host(ctx, &ok, all, sizeof(int)));
    for (int q = 0; q < N; q++) ok &= all[q];
    if (ok) {
        if (cudaMalloc((void**)&ctx->p2p_sync, 4 * CZK_P2P_MAX) != cudaSuccess) ok = 0;
        else cudaMemsetAsync(ctx->p2p_sync, 0, 4 * CZK_P2P_MAX, ctx->stream);
    }
    ctx->p2p_state = ok ? 1 : -1;
    return CZK_OK;
}

// grow a peer-addressable buffer: collective (every rank calls it with the same size at the same point of the protocol)
static int sh_p2p_reserve(czk_ctx* ctx, P2PBuf& b, size_t bytes) {
    if (bytes <= b.cap) return CZK_OK;
    const int N = ctx->nranks;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    // nobody may still be reading the old mapping: the host all-gathers below are the barrier
    for (int q = 0; q < N; q++)
        if (b.peer[q] && b.peer[q] != b.local) {
            cudaIpcCloseMemHandle(b.peer[q]);
            b.peer[q] = nullptr;
        }
    int token = 1, tokens[CZK_P2P_MAX];
    CZK_TRY(sh_allgather_host(ctx, &token, tokens, sizeof(int)));  // every rank has unmapped
    cudaFree(b.local);
    b.local = nullptr;
    b.cap = 0;
    const size_t want = bytes + bytes / 8;
    CUDA_TRY(ctx, cudaMalloc((void**)&b.local, want));
    // export, exchange, map - and agree on the outcome: if ANY rank could not map ANY peer, every rank drops the peer-memory
    // transport for this communicator (p2p_state = -1) and the caller goes through the NCCL exchange instead
    cudaIpcMemHandle_t mine, all[CZK_P2P_MAX];
    std::memset(&mine, 0, sizeof mine);
    int ok = cudaIpcGetMemHandle(&mine, b.local) == cudaSuccess;
    CZK_TRY(sh_allgather_host(ctx, &mine, all, sizeof mine));
    for (int q = 0; q < N; q++) {
        if (q == ctx->rank) {
            b.peer[q] = b.local;
            continue;
        }
        void* p = nullptr;
        if (ok && cudaIpcOpenMemHandle(&p, all[q], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) b.peer[q] = (uint32_t*)p;
        else ok = 0;
    }
    cudaGetLastError();  // a failed IPC call must not poison later error checks
    int oks[CZK_P2P_MAX] = {0};
    CZK_TRY(sh_allgather_host(ctx, &ok, oks, sizeof(int)));
    for (int q = 0; q < N; q++) ok &= oks[q];
    if (!ok) {
        for (int q = 0; q < N; q++)
            if (b.peer[q] && b.peer[q] != b.local) cudaIpcCloseMemHandle(b.peer[q]);
        for (int q = 0; q < N; q++) b.peer[q] = nullptr;
        CZK_TRY(sh_allgather_host(ctx, &token, tokens, sizeof(int)));  // every rank has unmapped before anybody frees
        cudaFree(b.local);
        b.local = nullptr;
        ctx->p2p_state = -1;
        return CZK_OK;
    }
    b.cap = want;
    return CZK_OK;
}

static int sh_barrier(czk_ctx* ctx) {
    ncclResult_t r = nccl_api().AllGather(ctx->p2p_sync + ctx->rank, ctx->p2p_sync, 4, ncclUint8, ctx->comm, ctx->stream);
    if (r != ncclSuccess) return fail(ctx, CZK_ERR_NCCL, std::string("ncclAllGather(barrier): ") + nccl_api().GetErrorString(r));
    return CZK_OK;
}

// The whole job over peer memory.  Same contract as sh_run_nccl: opened values in j.opened[0 .. total) on every rank,
// this rank's verdict in *j.flag.
static int sh_run_p2p(czk_ctx* ctx, const ShShape& s, ShJob& j) {
    const int N = ctx->nranks, me = ctx->rank;
    const bool spdz = s.scheme == CZK_SCHEME_SPDZ;
    const size_t bytes = s.padded * 32, slice_words = s.m * 8;
    CZK_TRY(sh_p2p_reserve(ctx, ctx->p2p_send, bytes));
    if (ctx->p2p_state == 1) CZK_TRY(sh_p2p_reserve(ctx, ctx->p2p_opened, bytes));
    if (ctx->p2p_state == 1 && spdz) CZK_TRY(sh_p2p_reserve(ctx, ctx->p2p_sigma, bytes));
    if (ctx->p2p_state != 1) return sh_run_nccl(ctx, s, j);  // the mapping failed somewhere: agreed by all ranks (see sh_p2p_reserve)
    j.send = ctx->p2p_send.local;
    j.opened = ctx->p2p_opened.local;
    j.sigma = spdz ? ctx->p2p_sigma.local : nullptr;
    CZK_TRY(sh_launch_pack(ctx, s, j));
    CZK_TRY(sh_barrier(ctx));  // every rank has packed
    {
        ShPtrs src{};
        ShDsts dst{};
        for (int p = 0; p < N; p++) {
            src.p[p] = ctx->p2p_send.peer[p] + (size_t)me * slice_words;    // slice `me` of rank p's send buffer (NVLink loads)
            dst.p[p] = ctx->p2p_opened.peer[p] + (size_t)me * slice_words;  // ... summed into every rank's opened buffer (NVLink stores)
        }
        k_sh_reduce_slices<<<sh_grid(s.m), 256, 0, ctx->stream>>>(dst, N, src, N, s.m); CZK_LAUNCHED();
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->link_bytes[0] += s.m * 32 * (uint64_t)(N - 1);  // stores into the peers
        ctx->link_bytes[1] += s.m * 32 * (uint64_t)(N - 1);  // loads from the peers
    }
    CZK_TRY(sh_barrier(ctx));  // every rank has written its slice into my opened buffer
    CZK_TRY(sh_launch_after_open(ctx, s, j));
    if (spdz) {
        CZK_TRY(sh_barrier(ctx));  // every rank's sigma is complete
        ShPtrs src{};
        for (int p = 0; p < N; p++) src.p[p] = ctx->p2p_sigma.peer[p] + (size_t)me * slice_words;
        k_sh_check_zero<<<sh_grid(s.m), 256, 0, ctx->stream>>>(src, N, s.m, j.flag); CZK_LAUNCHED();
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->link_bytes[1] += s.m * 32 * (uint64_t)(N - 1);
    }
    return CZK_OK;
}

// transport choice for this communicator (collective on first use)
static int sh_run(czk_ctx* ctx, const ShShape& s, ShJob& j) {
    if (ctx->nranks > 1) {
        CZK_TRY(sh_p2p_probe(ctx));
        if (ctx->p2p_state == 1) return sh_run_p2p(ctx, s, j);
    }
    return sh_run_nccl(ctx, s, j);
}

static int sh_reserve(czk_ctx* ctx, const ShShape& s, ShJob& j) {
    const size_t bytes = s.padded * 32;
    CZK_TRY(scratch_reserve(ctx, ctx->open_sx, bytes));
    j.send = (uint32_t*)ctx->open_sx.p;
    if (ctx->nranks > 1) {
        CZK_TRY(scratch_reserve(ctx, ctx->open_gather, bytes));
        CZK_TRY(scratch_reserve(ctx, ctx->open_oy, bytes));
        j.recv = (uint32_t*)ctx->open_gather.p;
        j.opened = (uint32_t*)ctx->open_oy.p;
    }
    if (s.scheme == CZK_SCHEME_SPDZ) {
        CZK_TRY(scratch_reserve(ctx, ctx->open_sigma, bytes));
        j.sigma = (uint32_t*)ctx->open_sigma.p;
    }
    return CZK_OK;
}

// One read-back per protocol call: every rank's verdict word is all-gathered (a failed check anywhere fails everywhere,
// like the reference's assert! at every party), then read with the stream synchronised once.
int sh_collect_flags(czk_ctx* ctx, const char* what) {
    uint32_t flags[SH_MAXP] = {0};
    const int N = ctx->nranks;
    if (N > 1) {
        CZK_TRY(scratch_reserve(ctx, ctx->open_d, 4 * SH_MAXP));
        ncclResult_t r = nccl_api().AllGather(ctx->flag, ctx->open_d.p, 4, ncclUint8, ctx->comm, ctx->stream);
        if (r != ncclSuccess) return fail(ctx, CZK_ERR_NCCL, std::string("ncclAllGather(flags): ") + nccl_api().GetErrorString(r));
        CUDA_TRY(ctx, cudaMemcpyAsync(flags, ctx->open_d.p, 4 * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        CUDA_TRY(ctx, cudaMemcpyAsync(flags, ctx->flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    uint32_t any = 0;
    for (int p = 0; p < N; p++) any |= flags[p];
    if (any) {
        cudaMemsetAsync(ctx->flag, 0, 4, ctx->stream);
        return fail(ctx, CZK_ERR_PROTOCOL, std::string(what) + ": SPDZ MAC check failed (spdz.rs:182 assert!(sum.is_zero()))");
    }
    return CZK_OK;
}

int czk_batch_open(czk_ctx* ctx, int scheme, const czk_vec* sh, const czk_vec* mac, czk_vec* out_pub, size_t n) {
    if (!ctx || !sh || !out_pub || n > sh->n || n > out_pub->n || (mac && n > mac->n))
        return fail(ctx, CZK_ERR_ARG, "czk_batch_open: range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (scheme == CZK_SCHEME_PLAIN) {
        if (out_pub->d != sh->d) CUDA_TRY(ctx, cudaMemcpyAsync(out_pub->d, sh->d, n * 32, cudaMemcpyDeviceToDevice, ctx->stream));
        return CZK_OK;
    }
    if (scheme != CZK_SCHEME_ADDITIVE && scheme != CZK_SCHEME_SPDZ) return fail(ctx, CZK_ERR_ARG, "czk_batch_open: scheme");
    if (scheme == CZK_SCHEME_SPDZ && !mac) return fail(ctx, CZK_ERR_ARG, "SPDZ open needs the MAC share vector");
    if (ctx->nranks > SH_MAXP) return fail(ctx, CZK_ERR_ARG, "additive / SPDZ shares: more than 16 parties");
    if (!n) return CZK_OK;
    const ShShape s = sh_shape(scheme, ctx->nranks, false, n);
    ShJob j;
    j.flag = ctx->flag;
    j.rank = ctx->rank;
    j.a_sh = (const uint32_t*)sh->d;
    j.a_mac = mac ? (const uint32_t*)mac->d : nullptr;
    sh_count_reference_open(ctx, scheme, n);
    CZK_TRY(sh_run(ctx, s, j));
    CUDA_TRY(ctx, cudaMemcpyAsync(out_pub->d, j.opened, n * 32, cudaMemcpyDeviceToDevice, ctx->stream));
    if (scheme == CZK_SCHEME_SPDZ) return sh_collect_flags(ctx, "czk_batch_open");
    return CZK_OK;
}

int czk_beaver_batch_mul(czk_ctx* ctx, int scheme, czk_vec* x_sh, czk_vec* x_mac, const czk_vec* y_sh, const czk_vec* y_mac,
                         size_t n) {
    if (!ctx || !x_sh || !y_sh || n > x_sh->n || n > y_sh->n) return fail(ctx, CZK_ERR_ARG, "czk_beaver_batch_mul: range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (scheme == CZK_SCHEME_PLAIN) return czk_vec_mul(ctx, x_sh, y_sh, n);
    // GszFieldShare::batch_mul (gsz20/mod.rs:309-315): king degree reduction, triple queued for the product check
    if (scheme == CZK_SCHEME_GSZ) return czk_gsz_batch_mul(ctx, x_sh, y_sh, n, 1);
    CZK_TRY(sh_beaver_mul_enqueue(ctx, scheme, x_sh, x_mac, y_sh, y_mac, n));
    if (scheme == CZK_SCHEME_SPDZ && n) return sh_collect_flags(ctx, "czk_beaver_batch_mul");
    return CZK_OK;
}

// The additive / SPDZ product, enqueued only: for SPDZ the verdict of the MAC check stays on the device (ctx->flag) until
// sh_collect_flags reads it.
int sh_beaver_mul_enqueue(czk_ctx* ctx, int scheme, czk_vec* x_sh, czk_vec* x_mac, const czk_vec* y_sh, const czk_vec* y_mac, size_t n) {
    if (!ctx || !x_sh || !y_sh || n > x_sh->n || n > y_sh->n) return fail(ctx, CZK_ERR_ARG, "czk_beaver_batch_mul: range");
    if (scheme != CZK_SCHEME_ADDITIVE && scheme != CZK_SCHEME_SPDZ) return fail(ctx, CZK_ERR_ARG, "czk_beaver_batch_mul: scheme");
    const bool spdz = scheme == CZK_SCHEME_SPDZ;
    if (spdz && (!x_mac || !y_mac || n > x_mac->n || n > y_mac->n)) return fail(ctx, CZK_ERR_ARG, "SPDZ product needs MAC vectors");
    if (ctx->nranks > SH_MAXP) return fail(ctx, CZK_ERR_ARG, "additive / SPDZ shares: more than 16 parties");
    if (!n) return CZK_OK;
    const ShShape s = sh_shape(scheme, ctx->nranks, true, n);
    ShJob j;
    j.flag = ctx->flag;
    j.rank = ctx->rank;
    j.a_sh = (const uint32_t*)x_sh->d;
    j.b_sh = (const uint32_t*)y_sh->d;
    j.a_mac = spdz ? (const uint32_t*)x_mac->d : nullptr;
    j.b_mac = spdz ? (const uint32_t*)y_mac->d : nullptr;
    j.out_sh = (uint32_t*)x_sh->d;
    j.out_mac = spdz ? (uint32_t*)x_mac->d : nullptr;
    sh_count_reference_open(ctx, scheme, n);  // the reference opens s + x and o + y one after the other
    sh_count_reference_open(ctx, scheme, n);
    return sh_run(ctx, s, j);
}

int czk_net_share_transport(const czk_ctx* ctx) { return ctx ? ctx->p2p_state : 0; }

int czk_net_link_bytes(const czk_ctx* ctx, uint64_t out[2]) {
    if (!ctx || !out) return CZK_ERR_ARG;
    out[0] = ctx->link_bytes[0];
    out[1] = ctx->link_bytes[1];
    return CZK_OK;
}

// ------------------------------------------------------------------------------------------ N parties on one GPU (diagnostics)
// The same kernels, the same slice geometry and the same per-party constants as sh_run_nccl, with the collectives replaced
// by direct addressing between the simulated parties' buffers: party q's reduce reads slice q of every party's send
// buffer and writes the sum into every party's `opened`.  This is what gives the N-party arithmetic a parity test that
// runs on one GPU (tests/test_gpu_shares.py); the transport itself is covered by the torchrun test.
struct ShSim {
    std::vector<ShJob> jobs;
    std::vector<void*> owned;
    uint32_t* flags = nullptr;
    ~ShSim() {
        for (void* p : owned) cudaFree(p);
    }
};
static int sh_sim_setup(czk_ctx* ctx, const ShShape& s, ShSim& sim) {
    const int N = s.parties;
    const bool spdz = s.scheme == CZK_SCHEME_SPDZ;
    sim.jobs.resize((size_t)N);
    auto grab = [&](size_t bytes, uint32_t** out) -> int {
        void* p = nullptr;
        CUDA_TRY(ctx, cudaMalloc(&p, bytes));
        sim.owned.push_back(p);
        *out = (uint32_t*)p;
        return CZK_OK;
    };
    CZK_TRY(grab(4 * (size_t)N, &sim.flags));
    CUDA_TRY(ctx, cudaMemsetAsync(sim.flags, 0, 4 * (size_t)N, ctx->stream));
    for (int q = 0; q < N; q++) {
        ShJob& j = sim.jobs[(size_t)q];
        j.rank = q;
        CZK_TRY(grab(s.padded * 32, &j.send));
        CZK_TRY(grab(s.padded * 32, &j.opened));
        if (spdz) CZK_TRY(grab(s.padded * 32, &j.sigma));
        j.flag = sim.flags + q;
    }
    return CZK_OK;
}
static int sh_sim_run(czk_ctx* ctx, const ShShape& s, ShSim& sim) {
    const int N = s.parties;
    const bool spdz = s.scheme == CZK_SCHEME_SPDZ;
    for (int q = 0; q < N; q++) CZK_TRY(sh_launch_pack(ctx, s, sim.jobs[(size_t)q]));
    for (int q = 0; q < N; q++) {  // party q reduces slice q and "all-gathers" it
        ShPtrs src{};
        ShDsts dst{};
        for (int p = 0; p < N; p++) {
            src.p[p] = sim.jobs[(size_t)p].send + (size_t)q * s.m * 8;
            dst.p[p] = sim.jobs[(size_t)p].opened + (size_t)q * s.m * 8;
        }
        k_sh_reduce_slices<<<sh_grid(s.m), 256, 0, ctx->stream>>>(dst, N, src, N, s.m); CZK_LAUNCHED();
        CUDA_TRY(ctx, cudaGetLastError());
    }
    for (int q = 0; q < N; q++) CZK_TRY(sh_launch_after_open(ctx, s, sim.jobs[(size_t)q]));
    if (spdz) {
        for (int q = 0; q < N; q++) {
            ShPtrs src{};
            for (int p = 0; p < N; p++) src.p[p] = sim.jobs[(size_t)p].sigma + (size_t)q * s.m * 8;
            k_sh_check_zero<<<sh_grid(s.m), 256, 0, ctx->stream>>>(src, N, s.m, sim.jobs[(size_t)q].flag); CZK_LAUNCHED();
            CUDA_TRY(ctx, cudaGetLastError());
        }
    }
    return CZK_OK;
}
static int sh_sim_flags(czk_ctx* ctx, const ShSim& sim, int N, uint32_t* flags_out) {
    std::vector<uint32_t> f((size_t)N, 0);
    CUDA_TRY(ctx, cudaMemcpyAsync(f.data(), sim.flags, 4 * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < N; q++) flags_out[q] = f[(size_t)q];
    return CZK_OK;
}

int czk_diag_sim_batch_open(czk_ctx* ctx, int scheme, int parties, const czk_vec* const* sh, const czk_vec* const* mac,
                            czk_vec* const* out_pub, size_t n, uint32_t* flags_out) {
    if (!ctx || !sh || !out_pub || !flags_out || parties < 1 || parties > SH_MAXP || !n ||
        (scheme != CZK_SCHEME_ADDITIVE && scheme != CZK_SCHEME_SPDZ) || (scheme == CZK_SCHEME_SPDZ && !mac))
        return fail(ctx, CZK_ERR_ARG, "czk_diag_sim_batch_open: argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const ShShape s = sh_shape(scheme, parties, false, n);
    ShSim sim;
    CZK_TRY(sh_sim_setup(ctx, s, sim));
    for (int q = 0; q < parties; q++) {
        if (!sh[q] || !out_pub[q] || sh[q]->n < n || out_pub[q]->n < n || (mac && (!mac[q] || mac[q]->n < n)))
            return fail(ctx, CZK_ERR_ARG, "czk_diag_sim_batch_open: vector");
        sim.jobs[(size_t)q].a_sh = (const uint32_t*)sh[q]->d;
        sim.jobs[(size_t)q].a_mac = mac ? (const uint32_t*)mac[q]->d : nullptr;
    }
    CZK_TRY(sh_sim_run(ctx, s, sim));
    for (int q = 0; q < parties; q++)
        CUDA_TRY(ctx, cudaMemcpyAsync(out_pub[q]->d, sim.jobs[(size_t)q].opened, n * 32, cudaMemcpyDeviceToDevice, ctx->stream));
    return sh_sim_flags(ctx, sim, parties, flags_out);
}

int czk_diag_sim_beaver_mul(czk_ctx* ctx, int scheme, int parties, czk_vec* const* x_sh, czk_vec* const* x_mac,
                            const czk_vec* const* y_sh, const czk_vec* const* y_mac, size_t n, uint32_t* flags_out) {
    const bool spdz = scheme == CZK_SCHEME_SPDZ;
    if (!ctx || !x_sh || !y_sh || !flags_out || parties < 1 || parties > SH_MAXP || !n || (scheme != CZK_SCHEME_ADDITIVE && !spdz) ||
        (spdz && (!x_mac || !y_mac)))
        return fail(ctx, CZK_ERR_ARG, "czk_diag_sim_beaver_mul: argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const ShShape s = sh_shape(scheme, parties, true, n);
    ShSim sim;
    CZK_TRY(sh_sim_setup(ctx, s, sim));
    for (int q = 0; q < parties; q++) {
        if (!x_sh[q] || !y_sh[q] || x_sh[q]->n < n || y_sh[q]->n < n || (spdz && (!x_mac[q] || !y_mac[q] || x_mac[q]->n < n || y_mac[q]->n < n)))
            return fail(ctx, CZK_ERR_ARG, "czk_diag_sim_beaver_mul: vector");
        ShJob& j = sim.jobs[(size_t)q];
        j.a_sh = (const uint32_t*)x_sh[q]->d;
        j.b_sh = (const uint32_t*)y_sh[q]->d;
        j.a_mac = spdz ? (const uint32_t*)x_mac[q]->d : nullptr;
        j.b_mac = spdz ? (const uint32_t*)y_mac[q]->d : nullptr;
        j.out_sh = (uint32_t*)x_sh[q]->d;
        j.out_mac = spdz ? (uint32_t*)x_mac[q]->d : nullptr;
    }
    CZK_TRY(sh_sim_run(ctx, s, sim));
    return sh_sim_flags(ctx, sim, parties, flags_out);
}
