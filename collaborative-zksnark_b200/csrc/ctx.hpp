// Internal (not installed) definitions shared by the translation units of libczk_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/czk.h"
#include "host_field.hpp"
#include "msm.cuh"

using namespace czk;
using namespace czk::host;

inline std::string& czk_tls_error() {
    static thread_local std::string e;
    return e;
}

// ------------------------------------------------------------------------------------------ NCCL (resolved at run time)
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
inline NcclApi& nccl_api() {
    static NcclApi api;
    if (api.handle || api.ok) return api;
    // if torch already loaded its bundled libnccl.so.2 the loader hands back that copy
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) return api;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
    api.Broadcast = (decltype(api.Broadcast))dlsym(api.handle, "ncclBroadcast");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
    api.Send = (decltype(api.Send))dlsym(api.handle, "ncclSend");
    api.Recv = (decltype(api.Recv))dlsym(api.handle, "ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))dlsym(api.handle, "ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.handle, "ncclGroupEnd");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.Broadcast && api.GetErrorString &&
             api.Send && api.Recv && api.GroupStart && api.GroupEnd;
    return api;
}

// ------------------------------------------------------------------------------------------ objects
struct czk_vec {
    uint64_t* d = nullptr;
    size_t n = 0;
};
struct czk_bases {
    int curve = 1;
    uint32_t* xy = nullptr;  // n * (24 | 48) words
    uint8_t* inf = nullptr;  // n bytes, or nullptr when no point is infinity
    size_t n = 0;
    uint32_t* table = nullptr;  // merged-window table: pre_w slabs of n affine points, slab w = 2^(pre_c w) * xy
    unsigned pre_c = 0, pre_w = 0;
};
struct Domain {
    int log_d = 0;
    uint32_t* tw = nullptr;
    uint32_t *g_lo = nullptr, *g_hi = nullptr, *gi_lo = nullptr, *gi_hi = nullptr, *g_hi_sinv = nullptr;
    uint32_t* g_sinv_br = nullptr;  // D^-1 g^bitrev(p) at position p: the fused iFFT -> coset FFT pair's scaling, built on first use
    int lo_log = 0;
    HFr size_inv, group_gen, group_gen_inv, generator_inv;
};
// MixedRadixEvaluationDomain of 3 * 2^log_m points: w = get_root_of_unity(3 M), tables w^j and w^-j (j < M)
struct MixedDomain {
    int log_m = 0;
    uint32_t *wpow = nullptr, *wipow = nullptr;
    HFr group_gen, group_gen_inv, size_inv, third_inv, zeta, zeta_inv;
};
// what the last GSZ group product check opened (czk_groth16_gsz_last_checks)
struct GszCheckOut {
    uint64_t group_x[4] = {0, 0, 0, 0};
    uint64_t group_yz[24] = {};
    uint8_t group_inf[2] = {0, 0};
};
struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
};

// GSZ20 (honest-majority Shamir) state of one party: the share domain of n parties and the queued product triples
struct GszTriple {
    uint64_t *x = nullptr, *y = nullptr, *z = nullptr;  // device copies, n elements each
    size_t n = 0;
};
struct GszState {
    int n = 0, t = 0;              // parties the tables below were built for, t = (n - 1) / 2
    uint32_t* winv_dev = nullptr;  // n Montgomery Fr: w^-k, w = get_root_of_unity(n)
    std::vector<HFr> winv_host;    // the same table for the host-side group opens
    HFr n_inv;
    std::vector<GszTriple> queue;  // field triples awaiting hadamard_check
    uint64_t king_computes = 0, opens = 0;
    uint64_t last_check[12] = {0};  // opened x, y, z of the last field product check
    Scratch dot_partial, one_elem, pad_x, pad_y, gather;
};

// One of the two MSM pipelines of a context: its own stream, workspace and pinned staging for the window sums, so that two
// MSMs can be in flight at once (the serial / latency-bound stretches of one - digit sort, finish walk, bucket reduction,
// host tail - run under the other's accumulation rounds).  Lane 0 also serves the synchronous entry points.
constexpr int CZK_MSM_SLOTS = 4;  // jobs that may wait uncollected on one lane
struct MsmSlot {
    void* pinned = nullptr;          // the job's window sums land here
    cudaEvent_t ev_done = nullptr;   // ... and this fires when they have
    cudaEvent_t ev_t[4] = {nullptr, nullptr, nullptr, nullptr};  // accumulate start/stop, whole MSM start/stop (device timing)
};
struct MsmLane {
    cudaStream_t stream = nullptr;
    MsmWorkspace ws;
    MsmSlot slots[CZK_MSM_SLOTS];
    cudaEvent_t ev_in = nullptr;  // recorded on the context stream: the lane's work starts after everything enqueued before it
    uint64_t enqueued = 0, collected = 0;  // jobs are collected in the order they were enqueued
};
// An MSM in flight (enqueued on a lane, not yet collected)
struct MsmJob {
    int lane = -1, curve = 1;
    uint64_t seq = 0;
    size_t n = 0;
    MsmConfig cfg;
};
constexpr int CZK_MSM_LANES = 2;

// A buffer of this rank that the other ranks of the box address directly (CUDA IPC over NVLink peer memory): `local` is this
// rank's allocation, peer[q] is rank q's allocation mapped into this process (peer[rank] == local).
constexpr int CZK_P2P_MAX = 16;
struct P2PBuf {
    uint32_t* local = nullptr;
    size_t cap = 0;
    uint32_t* peer[CZK_P2P_MAX] = {nullptr};
};

struct czk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int lane_priority = 0;
    std::string err;
    std::map<int, Domain> domains;
    std::map<int, MixedDomain> mixed_domains;
    MsmLane lanes[CZK_MSM_LANES];
    Scratch up_bases, up_inf, up_scalars, up_vec;  // staging for the host-pointer entry points
    Scratch open_gather, open_sigma, open_sx, open_oy, open_d, open_dm;
    Scratch mixed;  // the three de-interleaved subsequences of every vector of a mixed-radix batch
    uint32_t* flag = nullptr;
    double phases[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // host-side phase times of the last proof (czk_groth16_last_phases)
    GszCheckOut gsz_check;
    cudaEvent_t ev_phase[2] = {nullptr, nullptr};  // witness-map start / stop on the context stream (phase report)
    // network
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    uint64_t stats[5] = {0, 0, 0, 0, 0};  // mpc-net's own accounting (what the reference would count)
    uint64_t link_bytes[2] = {0, 0};      // bytes this rank actually sent / received over NVLink
    // share opens over peer memory (shares.cu): 0 = not probed yet, 1 = every rank can address every other, -1 = NCCL exchange
    int p2p_state = 0;
    P2PBuf p2p_send, p2p_opened, p2p_sigma;
    uint32_t* p2p_sync = nullptr;  // N words: the tiny all-gather that orders the kernels of different ranks
    GszState gsz;
    // kernel timing of the MSM (CUDA events on the launching stream), per curve: [0] G1, [1] G2
    double acc_ms[2] = {0, 0}, msm_ms[2] = {0, 0}, acc_terms[2] = {0, 0}, acc_entries[2] = {0, 0};
    uint64_t acc_launches[2] = {0, 0};
};

// asynchronous MSM over resident bases (api.cu): enqueue on lane 0 / 1, collect in enqueue order per lane
int msm_bases_enqueue(czk_ctx* ctx, int lane, const czk_bases* b, size_t base_off, const czk_vec* sc, size_t sc_off,
                      int scalars_montgomery, size_t n, MsmJob* job, bool reuse_plan = false);
int msm_collect(czk_ctx* ctx, MsmJob* job, uint64_t* out_xyz, double* device_ms);

// shares.cu: the Beaver product without the final verdict read-back, and the read-back itself (one stream synchronisation)
int sh_beaver_mul_enqueue(czk_ctx* ctx, int scheme, czk_vec* x_sh, czk_vec* x_mac, const czk_vec* y_sh, const czk_vec* y_mac, size_t n);
int sh_collect_flags(czk_ctx* ctx, const char* what);
void sh_p2p_release(czk_ctx* ctx);  // unmap / free the peer-addressable buffers (before the communicator goes away)

inline int fail(czk_ctx* ctx, int code, const std::string& msg) {
    czk_tls_error() = msg;
    if (ctx) ctx->err = msg;
    return code;
}
#define CUDA_TRY(ctx, expr)                                                                                  \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            return fail(ctx, CZK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));              \
    } while (0)
#define CZK_TRY(expr)            \
    do {                         \
        int _s = (expr);         \
        if (_s != CZK_OK) return _s; \
    } while (0)

inline int scratch_reserve(czk_ctx* ctx, Scratch& s, size_t bytes) {
    if (bytes <= s.cap) return CZK_OK;
    if (s.p) CUDA_TRY(ctx, cudaFree(s.p));
    s.p = nullptr;
    s.cap = 0;
    size_t want = bytes + bytes / 8;
    CUDA_TRY(ctx, cudaMalloc(&s.p, want));
    s.cap = want;
    return CZK_OK;
}

