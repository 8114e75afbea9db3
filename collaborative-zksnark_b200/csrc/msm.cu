// Device kernels of the bucket-method MSM (see msm.cuh for the pipeline and the reference lines replaced).
#include "msm.cuh"

#include <cstdlib>
#include <cstring>

#include "launch_count.hpp"
#include "msm_io.cuh"

namespace czk {

size_t msm_point_words(int curve) { return curve == 1 ? 48 : 96; }

// ------------------------------------------------------------------ 1. prepare
__global__ void k_msm_prepare(const uint32_t* __restrict__ scalars_in, const uint8_t* __restrict__ inf,
                              uint32_t* __restrict__ scalars_out, uint32_t* __restrict__ hist, size_t n, unsigned c,
                              unsigned nwin, unsigned nb, int mont, int merged) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* q = reinterpret_cast<const uint4*>(scalars_in) + 2 * i;
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fr s;
    s.l[0] = a.x; s.l[1] = a.y; s.l[2] = a.z; s.l[3] = a.w;
    s.l[4] = b.x; s.l[5] = b.y; s.l[6] = b.z; s.l[7] = b.w;
    if (mont) s = Fr::from_mont(s);
    if (inf && inf[i]) s = Fr::zero();
    uint4* o = reinterpret_cast<uint4*>(scalars_out) + 2 * i;
    o[0] = make_uint4(s.l[0], s.l[1], s.l[2], s.l[3]);
    o[1] = make_uint4(s.l[4], s.l[5], s.l[6], s.l[7]);
    uint32_t sl[8];
#pragma unroll
    for (int k = 0; k < 8; k++) sl[k] = s.l[k];
    DigitCursor cur;
    for (unsigned w = 0; w < nwin; w++) {
        int32_t d = cur.next(sl, c, w);
        if (d != 0) {
            uint32_t mag = (uint32_t)(d < 0 ? -d : d);
            atomicAdd(&hist[(merged ? 0 : (size_t)w * nb) + (mag - 1)], 1u);
        }
    }
}

// ------------------------------------------------------------------ 2. exclusive scan (one block, coalesced tiles)
__global__ void __launch_bounds__(1024) k_exclusive_scan(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (size_t tile = 0; tile < n; tile += 4096) {
        size_t i0 = tile + (size_t)tid * 4;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = (i0 + k < n) ? in[i0 + k] : 0u;
        uint32_t tsum = v[0] + v[1] + v[2] + v[3];
        uint32_t inc = tsum;  // inclusive scan of the per-thread sums within the warp
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = warp_sums[lane];
            uint32_t winc = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, winc, off);
                if (lane >= off) winc += t;
            }
            warp_sums[lane] = winc - w;  // exclusive prefix of the warp totals
        }
        __syncthreads();
        uint32_t base = carry_s + warp_sums[wid] + (inc - tsum);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i0 + k < n) out[i0 + k] = base;
            base += v[k];
        }
        __syncthreads();
        if (tid == 1023) carry_s = base;
        __syncthreads();
    }
}

// ------------------------------------------------------------------ 3. scatter
__global__ void k_msm_scatter(const uint32_t* __restrict__ scalars, uint32_t* __restrict__ cursor,
                              uint32_t* __restrict__ sorted, size_t n, unsigned c, unsigned nwin, unsigned nb, int merged,
                              size_t stride, size_t off) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* q = reinterpret_cast<const uint4*>(scalars) + 2 * i;
    uint4 a = q[0], b = q[1];
    uint32_t sl[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    DigitCursor cur;
    for (unsigned w = 0; w < nwin; w++) {
        int32_t d = cur.next(sl, c, w);
        if (d != 0) {
            uint32_t mag = (uint32_t)(d < 0 ? -d : d);
            uint32_t pos = atomicAdd(&cursor[(merged ? 0 : (size_t)w * nb) + (mag - 1)], 1u);
            // merged: the entry addresses slab w of the precomputed table, 2^(c w) * P_i
            uint32_t idx = merged ? (uint32_t)((size_t)w * stride + off + i) : (uint32_t)i;
            sorted[pos] = idx | (d < 0 ? 0x80000000u : 0u);
        }
    }
}

// ------------------------------------------------------------------ 4. accumulate
// A bucket's sorted run is cut into segments of at most `seg` points; one thread sums one segment.
// With uniform digits every bucket is a single segment (seg >= the mean bucket load) and the result
// goes straight to buckets[]; heavier buckets - a degenerate top window, repeated scalars - are split
// so no thread ever walks more than `seg` points, and a second kernel folds the segment sums.
__global__ void k_msm_seg_counts(const uint32_t* __restrict__ hist, uint32_t* __restrict__ segcnt, size_t total, uint32_t seg,
                                 uint32_t* __restrict__ maxlen) {
    size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t c = id < total ? hist[id] : 0u;
    if (id < total) segcnt[id] = c ? (c + seg - 1) / seg : 1u;  // empty buckets keep one (empty) segment so they get zeroed
    // longest bucket: the number of halving rounds of the batched-affine path
    uint32_t m = __reduce_max_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(maxlen, m);
}

// item descriptors: one per segment, found by binary search once (fully parallel) so the accumulation loop
// never searches.  desc = {first sorted entry, length, bucket, 1 if the bucket has several segments}
__global__ void k_msm_build_items(const uint32_t* __restrict__ ends, const uint32_t* __restrict__ hist,
                                  const uint32_t* __restrict__ segoff, const uint32_t* __restrict__ segcnt,
                                  uint4* __restrict__ items, uint32_t* __restrict__ nitems_out, uint32_t* __restrict__ heavy,
                                  size_t total, size_t max_items, uint32_t seg, uint32_t heavy_len) {
    size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t nitems = (size_t)segoff[total - 1] + segcnt[total - 1];
    if (id == 0) {
        nitems_out[0] = (uint32_t)nitems;
        nitems_out[1] = 0;  // work-queue head   ([2] = heavy count, zeroed by the host before this launch)
    }
    if (id >= max_items || id >= nitems) return;
    size_t lo = 0, hi = total - 1;
    while (lo < hi) {  // last bucket b with segoff[b] <= id
        size_t mid = (lo + hi + 1) >> 1;
        if (segoff[mid] <= id) lo = mid;
        else hi = mid - 1;
    }
    const uint32_t b = (uint32_t)lo;
    const uint32_t k = (uint32_t)(id - segoff[b]);
    const uint32_t cnt = hist[b];
    uint32_t len = cnt > k * seg ? cnt - k * seg : 0;
    if (len > seg) len = seg;
    // items much longer than the mean are listed separately and served first (longest-processing-time-first):
    // wherever they sit in bucket order - the windowed form's sparse top window, the merged form's low buckets -
    // they must not be the last thing a lane picks up
    uint32_t is_heavy = len >= heavy_len ? 2u : 0u;
    if (is_heavy) heavy[atomicAdd(&nitems_out[2], 1u)] = (uint32_t)id;
    items[id] = make_uint4(ends[b] - cnt + k * seg, len, b, (segcnt[b] > 1 ? 1u : 0u) | is_heavy);
}

// Lane-level dynamic scheduling: every lane of a resident warp owns one item at a time and adds ONE point per
// loop trip; a lane whose item is exhausted stores its sum and pulls the next item from a global queue inside
// the same trip.  Bucket loads differ (Poisson around N/2^(c-1)), but no lane waits for a longer neighbour:
// the warp only idles at the very end of the kernel.  The next point is fetched while the current one is added.
template <class F, bool PREFETCH, int MINB>
__global__ void __launch_bounds__(128, MINB) k_msm_accumulate(const uint32_t* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                                         const uint4* __restrict__ items, uint32_t* __restrict__ queue,
                                                         const uint32_t* __restrict__ heavy, uint32_t* __restrict__ buckets,
                                                         uint32_t* __restrict__ segsum, const uint32_t* __restrict__ gate,
                                                         const unsigned bstride) {
    constexpr int W = FieldIO<F>::W;
    if (gate && *gate == 0) return;  // the batched-affine path already produced the buckets
    const uint32_t nitems = queue[0], nheavy = queue[2];
    const unsigned lane = threadIdx.x & 31;
    XYZZ<F> acc = XYZZ<F>::infinity();
    uint32_t pos = 0, remaining = 0, item = 0xffffffffu, bucket = 0, multi = 0;
    bool alive = true;
    F nx, ny;            // prefetched point of the current trip
    uint32_t nsign = 0;
    while (true) {
        if (remaining == 0 && alive) {
            if (item != 0xffffffffu) {
                if (multi) store_point<F>(segsum + (size_t)item * (4 * W), acc);
                else store_point<F>(buckets + (size_t)bucket * (4 * W), acc);
            }
            // warp-aggregated fetch: one atomic for all lanes that need work
            unsigned need = __activemask();
            unsigned leader = __ffs(need) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(&queue[1], (uint32_t)__popc(need));
            base = __shfl_sync(need, base, leader);
            uint32_t q = base + __popc(need & ((1u << lane) - 1));
            if (q < nheavy + nitems) {
                // queue = the heavy items first, then every item in bucket order (heavy ones skipped there)
                bool from_heavy = q < nheavy;
                item = from_heavy ? heavy[q] : q - nheavy;
                uint4 d = items[item];
                pos = d.x;
                remaining = d.y;
                bucket = d.z;
                multi = d.w & 1u;
                acc = XYZZ<F>::infinity();
                if (!from_heavy && (d.w & 2u)) {  // already taken from the heavy list: nothing to do, nothing to store
                    remaining = 0;
                    item = 0xffffffffu;
                }
                if (PREFETCH && remaining) {
                    uint32_t e = sorted[pos];
                    const uint32_t* p = bases + (size_t)(e & 0x7fffffffu) * bstride;
                    nx = FieldIO<F>::load(p);
                    ny = FieldIO<F>::load(p + W);
                    nsign = e >> 31;
                }
            } else {
                alive = false;
                item = 0xffffffffu;
            }
        }
        if (!__any_sync(0xffffffffu, alive)) break;
        if (alive && remaining) {
            F x, y;
            uint32_t sg;
            if (PREFETCH) {
                x = nx;
                y = ny;
                sg = nsign;
            } else {  // G2: the prefetch registers would spill; load in place
                uint32_t e = sorted[pos];
                const uint32_t* p = bases + (size_t)(e & 0x7fffffffu) * bstride;
                x = FieldIO<F>::load(p);
                y = FieldIO<F>::load(p + W);
                sg = e >> 31;
            }
            remaining--;
            pos++;
            if (PREFETCH && remaining) {  // issue the next point's loads before the long addition
                uint32_t e = sorted[pos];
                const uint32_t* p = bases + (size_t)(e & 0x7fffffffu) * bstride;
                nx = FieldIO<F>::load(p);
                ny = FieldIO<F>::load(p + W);
                nsign = e >> 31;
            }
            if (sg) y = F::neg(y);
            acc.add_affine(x, y);
        }
    }
}

// fold the segment sums of the (rare) multi-segment buckets
template <class F>
__global__ void __launch_bounds__(128) k_msm_fold_segments(const uint32_t* __restrict__ segoff, const uint32_t* __restrict__ segcnt,
                                                            const uint32_t* __restrict__ segsum, uint32_t* __restrict__ buckets,
                                                            size_t total, const uint32_t* __restrict__ gate) {
    constexpr int W = FieldIO<F>::W;
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= total || (gate && *gate == 0)) return;
    uint32_t n = segcnt[b];
    if (n <= 1) return;
    size_t off = segoff[b];
    XYZZ<F> acc = XYZZ<F>::infinity();
    for (uint32_t k = 0; k < n; k++) acc.add(load_point<F>(segsum + (off + k) * (4 * W)));
    store_point<F>(buckets + b * (4 * W), acc);
}

// ------------------------------------------------------------------ 5a. chunked running sums
// thread t of window w owns buckets [lo, lo+chunk): emits  sum_b (b+1) B_b  over its chunk
template <class F>
__global__ void __launch_bounds__(128) k_msm_reduce_chunks(const uint32_t* __restrict__ buckets, uint32_t* __restrict__ partial,
                                                            unsigned nb, unsigned chunk, unsigned nwin) {
    constexpr int W = FieldIO<F>::W;
    unsigned nchunks = nb / chunk;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nwin * nchunks) return;
    unsigned w = (unsigned)(t / nchunks), ci = (unsigned)(t % nchunks);
    unsigned lo = ci * chunk;
    XYZZ<F> S = XYZZ<F>::infinity(), T = XYZZ<F>::infinity();
    for (unsigned k = chunk; k-- > 0;) {
        XYZZ<F> B = load_point<F>(buckets + ((size_t)w * nb + lo + k) * (4 * W));
        S.add(B);
        T.add(S);
    }
    if (lo) {
        // lo * S, MSB-first double and add
        XYZZ<F> R = XYZZ<F>::infinity();
        int top = 31 - __clz(lo);
        for (int bit = top; bit >= 0; bit--) {
            R = XYZZ<F>::dbl(R);
            if ((lo >> bit) & 1) R.add(S);
        }
        T.add(R);
    }
    store_point<F>(partial + t * (4 * W), T);
}

// ------------------------------------------------------------------ 5a'. bit-sliced subset sums (merged windows)
// sum_b (b+1) B_b = sum_j 2^j S_j with S_j = the sum of the buckets whose weight b+1 has bit j set: c plain sums of
// nb/2 points each (plus the single bucket of weight nb) - no per-thread scalar multiple, no running sum, every
// thread adds exactly members/T points.  Thread t of slice j sums the members k = t, t+T, ... of slice j; the k-th
// member is the weight v = (hi << (j+1)) | (1 << j) | lo with lo = k mod 2^j, hi = k div 2^j.
template <class F>
__global__ void __launch_bounds__(128) k_msm_bit_sums(const uint32_t* __restrict__ buckets, uint32_t* __restrict__ partial,
                                                       unsigned nb, unsigned T) {
    constexpr int W = FieldIO<F>::W;
    const unsigned j = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    // members of slice j among the weights 1..nb: nb/2 for j < log2(nb), one (v = nb) for j = log2(nb)
    const unsigned top = 31 - __clz(nb);
    const unsigned members = j < top ? nb / 2 : 1;
    XYZZ<F> acc = XYZZ<F>::infinity();
    for (unsigned k = t; k < members; k += T) {
        unsigned v = j < top ? ((((k >> j) << 1) | 1u) << j) | (k & ((1u << j) - 1)) : nb;
        acc.add(load_point<F>(buckets + (size_t)(v - 1) * (4 * W)));
    }
    store_point<F>(partial + ((size_t)j * T + t) * (4 * W), acc);
}

// ------------------------------------------------------------------ 5a''. the same slice sums through a two-level grid
// Write the weight v in [0, nb) as v = hi * L + lo (L = 2^lo_bits, H = nb / L; v = 0 has no bucket).  With the row sums
// R_hi = sum_lo B_v and the column sums C_lo = sum_hi B_v, slice j < lo_bits is the sum of the C_lo whose index has bit j
// set and slice lo_bits + j the sum of the R_hi whose index has bit j set; the top slice is the single bucket of weight
// nb.  Two adds per bucket instead of (c - 1) / 2: the bit-slice kernel above is throughput bound on G2.
// Block b < H sums row b, block H + b sums column b; GRID_THREADS threads stride over the L (or H) members, then a tree.
constexpr int GRID_THREADS = 64;
template <class F>
__device__ __forceinline__ XYZZ<F> block_tree_sum(XYZZ<F> acc, XYZZ<F>* sm, unsigned tid) {
    sm[tid] = acc;
    __syncthreads();
    for (unsigned s = GRID_THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) {
            XYZZ<F> a = sm[tid];
            a.add(sm[tid + s]);
            sm[tid] = a;
        }
        __syncthreads();
    }
    return sm[0];
}
template <class F>
__global__ void __launch_bounds__(GRID_THREADS) k_msm_grid_sums(const uint32_t* __restrict__ buckets, uint32_t* __restrict__ grid,
                                                                 unsigned lo_bits, unsigned hi_bits) {
    constexpr int W = FieldIO<F>::W;
    __shared__ XYZZ<F> sm[GRID_THREADS];
    const unsigned L = 1u << lo_bits, H = 1u << hi_bits, b = blockIdx.x, tid = threadIdx.x;
    const bool row = b < H;
    const unsigned fixed = row ? b : b - H, count = row ? L : H;
    XYZZ<F> acc = XYZZ<F>::infinity();
    for (unsigned k = tid; k < count; k += GRID_THREADS) {
        unsigned v = row ? fixed * L + k : k * L + fixed;
        if (v) acc.add(load_point<F>(buckets + (size_t)(v - 1) * (4 * W)));
    }
    acc = block_tree_sum<F>(acc, sm, tid);
    if (tid == 0) store_point<F>(grid + (size_t)b * (4 * W), acc);
}
// block j: slice j from the grid sums (grid[0..H) rows, grid[H..H+L) columns); the top slice copies bucket nb
template <class F>
__global__ void __launch_bounds__(GRID_THREADS) k_msm_grid_slices(const uint32_t* __restrict__ grid, const uint32_t* __restrict__ buckets,
                                                                   uint32_t* __restrict__ out, unsigned lo_bits, unsigned hi_bits) {
    constexpr int W = FieldIO<F>::W;
    __shared__ XYZZ<F> sm[GRID_THREADS];
    const unsigned L = 1u << lo_bits, H = 1u << hi_bits, j = blockIdx.x, tid = threadIdx.x;
    XYZZ<F> acc = XYZZ<F>::infinity();
    if (j == lo_bits + hi_bits) {
        if (tid == 0) acc = load_point<F>(buckets + (size_t)(H * L - 1) * (4 * W));
    } else {
        const bool col = j < lo_bits;
        const unsigned bit = col ? j : j - lo_bits, count = (col ? L : H) / 2;
        const uint32_t* src = col ? grid + (size_t)H * (4 * W) : grid;
        for (unsigned k = tid; k < count; k += GRID_THREADS) {
            unsigned idx = ((((k >> bit) << 1) | 1u) << bit) | (k & ((1u << bit) - 1));
            acc.add(load_point<F>(src + (size_t)idx * (4 * W)));
        }
    }
    acc = block_tree_sum<F>(acc, sm, tid);
    if (tid == 0) store_point<F>(out + (size_t)j * (4 * W), acc);
}

// ------------------------------------------------------------------ 5b. tree sums of the chunk results
// block b sums in[b*count .. (b+1)*count) -> out[b]; run twice (chunks -> groups -> window) so that the first
// level has windows*groups blocks instead of one block per window
constexpr int WINSUM_THREADS = 64;
constexpr int WINSUM_GROUPS = 32;
template <class F>
__global__ void __launch_bounds__(WINSUM_THREADS) k_msm_block_sum(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                   unsigned count) {
    constexpr int W = FieldIO<F>::W;
    __shared__ XYZZ<F> sm[WINSUM_THREADS];
    unsigned b = blockIdx.x, tid = threadIdx.x;
    XYZZ<F> acc = XYZZ<F>::infinity();
    for (unsigned k = tid; k < count; k += WINSUM_THREADS) acc.add(load_point<F>(in + ((size_t)b * count + k) * (4 * W)));
    sm[tid] = acc;
    __syncthreads();
    for (unsigned s = WINSUM_THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) {
            XYZZ<F> a = sm[tid];
            a.add(sm[tid + s]);
            sm[tid] = a;
        }
        __syncthreads();
    }
    if (tid == 0) store_point<F>(out + (size_t)b * (4 * W), sm[0]);
}

template <class F>
static cudaError_t msm_run_t(const uint32_t* bases, const uint8_t* inf, const uint32_t* scalars, bool mont, size_t n,
                             const MsmConfig& cfg, MsmWorkspace& ws, cudaStream_t st, bool reuse_plan) {
    size_t total = (size_t)cfg.bwin * cfg.nb;
    const unsigned bstride = cfg.base_stride ? cfg.base_stride : 2u * FieldIO<F>::W;
    cudaError_t e;
    if (ws.ev[2]) cudaEventRecord(ws.ev[2], st);
    if (!reuse_plan) {  // steps 1-3: digits, histogram, scan, scatter
        if ((e = cudaMemsetAsync(ws.hist, 0, total * sizeof(uint32_t), st)) != cudaSuccess) return e;
        if (n) {
            unsigned blocks = (unsigned)((n + 255) / 256);
            k_msm_prepare<<<blocks, 256, 0, st>>>(scalars, inf, ws.scalars, ws.hist, n, cfg.c, cfg.nwin, cfg.nb, mont ? 1 : 0, (int)cfg.merged); CZK_LAUNCHED();
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
        }
        k_exclusive_scan<<<1, 1024, 0, st>>>(ws.hist, ws.offsets, total); CZK_LAUNCHED();
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (n) {
            unsigned blocks = (unsigned)((n + 255) / 256);
            k_msm_scatter<<<blocks, 256, 0, st>>>(ws.scalars, ws.offsets, ws.sorted, n, cfg.c, cfg.nwin, cfg.nb, (int)cfg.merged, cfg.table_stride,
                                                    cfg.table_off); CZK_LAUNCHED();
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
        }
    }
    // segment length: at least the mean load of a busy bucket, and ~sqrt(n) so that neither the per-segment
    // walk nor the fold over segments can exceed O(sqrt(n)) serial additions whatever the digits are
    // item length: (a) at most ~sqrt(n), so neither a segment walk nor the fold over a bucket's segments can exceed
    // O(sqrt n) serial additions whatever the digits are; (b) short enough that every lane of the resident grid gets
    // several items (>= 8), or the last items of the queue run on a mostly idle machine
    uint32_t seg = 128;
    while ((size_t)seg * seg < n) seg <<= 1;
    {
        size_t lanes = (size_t)ws.sm_count * 256;
        size_t fine = (n * cfg.nwin) / (lanes * 8);
        if (fine < 32) fine = 32;
        if (fine < seg) seg = (uint32_t)fine;
    }
    size_t max_items = total + (n * cfg.nwin) / seg + 1;
    if (max_items > ws.cap_items) return cudaErrorInvalidValue;
    // queue words: 0 item count, 1 queue head, 2 heavy count, 3 "affine path gave up" flag, 4 longest bucket
    if (reuse_plan) {  // the items and their counts stand; rewind the queue head and clear the flag
        if ((e = cudaMemsetAsync(ws.queue + 1, 0, 4, st)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(ws.queue + 3, 0, 4, st)) != cudaSuccess) return e;
    } else {
        if ((e = cudaMemsetAsync(ws.queue, 0, 32, st)) != cudaSuccess) return e;
        k_msm_seg_counts<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ws.hist, ws.segcnt, total, seg, ws.queue + 4); CZK_LAUNCHED();
        k_exclusive_scan<<<1, 1024, 0, st>>>(ws.segcnt, ws.segoff, total); CZK_LAUNCHED();
        uint32_t heavy_len = (uint32_t)(2 * ((n * cfg.nwin) / total + 1) + 16);
        k_msm_build_items<<<(unsigned)((max_items + 255) / 256), 256, 0, st>>>(ws.offsets, ws.hist, ws.segoff, ws.segcnt, (uint4*)ws.items,
                                                                                ws.queue, ws.heavy, total, max_items, seg, heavy_len); CZK_LAUNCHED();
    }
    if (ws.ev[0]) cudaEventRecord(ws.ev[0], st);
    const uint32_t* gate = nullptr;
    // worth it only when buckets are long (merged windows over a large base set): the rounds' fixed latencies (one
    // block-wide inversion each) must be paid back by the multiplications saved (measured, whole MSM, with the rounds
    // decided on the device: 2^17 G1 terms 1.83 ms walk vs 1.75 tree, 2^18 3.09 vs 2.60, 2^19 5.04 vs 3.88; 2^18 G2 terms
    // 11.2 vs 6.9); CZK_BAT_MIN_ENTRIES / CZK_BAT_MIN_LOAD override for A/B runs
    static const size_t bat_min_entries = [] {
        const char* e = getenv("CZK_BAT_MIN_ENTRIES");
        return e ? (size_t)atoll(e) : (size_t)2 << 20;  // 2^18 terms x 15 windows and up
    }();
    static const size_t bat_min_load = [] {
        const char* e = getenv("CZK_BAT_MIN_LOAD");
        return e ? (size_t)atoll(e) : (size_t)24;
    }();
    if (ws.batched && n && (ws.batched_always || (n * cfg.nwin >= bat_min_entries && n * cfg.nwin >= total * bat_min_load))) {
        // tree of batched affine additions (msm_batched.cu).  The longest bucket (queue word 4, written by k_msm_seg_counts)
        // decides the number of halving rounds on the device: nothing is read back, the whole MSM is enqueued in one go
        if ((e = msm_batched_accumulate(FieldIO<F>::W == 12 ? 1 : 2, bases, bstride, ws.sorted, ws.offsets, ws.hist, total, n * cfg.nwin, ws.queue + 4,
                                        ws.bat_a, ws.bat_b, ws.bat_prefix, ws.buckets, ws.queue + 3, ws.sm_count, st)) != cudaSuccess)
            return e;
        gate = ws.queue + 3;
    }
    {
        // resident grid: the queue feeds lanes, so launch what the machine holds (2 blocks of 128 per SM at this
        // register budget) and no more; small problems launch fewer blocks
        size_t want = (max_items + 127) / 128;
        // resident grid: 2 blocks of 128 per SM is what the register budget allows; more (with fewer registers or
        // out-of-line products) measured 3-8 % slower
        size_t cap = (size_t)ws.sm_count * 2;
        unsigned blocks = (unsigned)(want < cap ? want : cap);
        k_msm_accumulate<F, (FieldIO<F>::W == 12), 2><<<blocks, 128, 0, st>>>(bases, ws.sorted, (const uint4*)ws.items, ws.queue, ws.heavy,
                                                                               ws.buckets, ws.segsum, gate, bstride);
        CZK_LAUNCHED();
    }
    if (ws.ev[1]) cudaEventRecord(ws.ev[1], st);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    k_msm_fold_segments<F><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(ws.segoff, ws.segcnt, ws.segsum, ws.buckets, total, gate); CZK_LAUNCHED();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    unsigned nchunks = cfg.nb / cfg.chunk;
    size_t rthreads = (size_t)cfg.bwin * nchunks;
    unsigned nsums = cfg.bwin;  // points left in winsum for the host tail
    if (msm_uses_bit_sums(cfg)) {
        // merged windows: c bit-slice sums (the host tail recombines them with c - 1 doublings)
        nsums = cfg.c;
        static const bool use_grid = [] { const char* e = getenv("CZK_MSM_REDUCE"); return !(e && !strcmp(e, "bits")); }();
        if (use_grid) {
            const unsigned top = 31 - __builtin_clz(cfg.nb), lo_bits = top / 2, hi_bits = top - lo_bits;
            k_msm_grid_sums<F><<<(1u << lo_bits) + (1u << hi_bits), GRID_THREADS, 0, st>>>(ws.buckets, ws.partial, lo_bits, hi_bits); CZK_LAUNCHED();
            k_msm_grid_slices<F><<<nsums, GRID_THREADS, 0, st>>>(ws.partial, ws.buckets, ws.winsum, lo_bits, hi_bits); CZK_LAUNCHED();
            if (ws.ev[3]) cudaEventRecord(ws.ev[3], st);
            return cudaGetLastError();
        }
        nchunks = msm_bitsum_threads(cfg);
        rthreads = (size_t)nsums * nchunks;
        k_msm_bit_sums<F><<<dim3(nchunks / 128, nsums), 128, 0, st>>>(ws.buckets, ws.partial, cfg.nb, nchunks); CZK_LAUNCHED();
    } else {
        k_msm_reduce_chunks<F><<<(unsigned)((rthreads + 127) / 128), 128, 0, st>>>(ws.buckets, ws.partial, cfg.nb, cfg.chunk, cfg.bwin); CZK_LAUNCHED();
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (nchunks >= 2 * WINSUM_GROUPS) {
        // partial[nwin*nchunks] -> segsum scratch [nwin*GROUPS] -> winsum[nwin]   (segsum is free again after the fold)
        uint32_t* mid = ws.partial + rthreads * msm_point_words(FieldIO<F>::W == 12 ? 1 : 2);
        k_msm_block_sum<F><<<nsums * WINSUM_GROUPS, WINSUM_THREADS, 0, st>>>(ws.partial, mid, nchunks / WINSUM_GROUPS); CZK_LAUNCHED();
        k_msm_block_sum<F><<<nsums, WINSUM_THREADS, 0, st>>>(mid, ws.winsum, WINSUM_GROUPS); CZK_LAUNCHED();
    } else {
        k_msm_block_sum<F><<<nsums, WINSUM_THREADS, 0, st>>>(ws.partial, ws.winsum, nchunks); CZK_LAUNCHED();
    }
    if (ws.ev[3]) cudaEventRecord(ws.ev[3], st);
    return cudaGetLastError();
}

cudaError_t msm_run(int curve, const uint32_t* bases, const uint8_t* inf, const uint32_t* scalars, bool scalars_mont,
                    size_t n, const MsmConfig& cfg, MsmWorkspace& ws, cudaStream_t st, bool reuse_plan) {
    if (curve == 1) return msm_run_t<Fq>(bases, inf, scalars, scalars_mont, n, cfg, ws, st, reuse_plan);
    return msm_run_t<Fq2>(bases, inf, scalars, scalars_mont, n, cfg, ws, st, reuse_plan);
}

__global__ void k_flags_differ(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n, uint32_t* __restrict__ differ) {
    bool d = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d |= a[i] != b[i];
    if (__any_sync(0xffffffffu, d) && (threadIdx.x & 31) == 0) atomicOr(differ, 1u);
}
cudaError_t msm_flags_differ(const uint8_t* a, const uint8_t* b, size_t n, uint32_t* differ, cudaStream_t st) {
    if (!n) return cudaSuccess;
    size_t blocks = (n + 255) / 256;
    k_flags_differ<<<(unsigned)(blocks < 1024 ? blocks : 1024), 256, 0, st>>>(a, b, n, differ); CZK_LAUNCHED();
    return cudaGetLastError();
}

// ------------------------------------------------------------------ merged-window table
// thread i: slab w holds 2^(c w) * P_i.  One doubling chain per base; every slab entry is normalised to affine
// (Fermat inversion): a one-off cost per CRS query, amortised over every proof made with the key.
template <class F>
__global__ void __launch_bounds__(128) k_msm_precompute(uint32_t* __restrict__ table, const unsigned ts, const uint32_t* __restrict__ bases,
                                                         size_t n, unsigned c, unsigned nwin) {
    constexpr int W = FieldIO<F>::W;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = FieldIO<F>::load(bases + i * (2 * W)), y = FieldIO<F>::load(bases + i * (2 * W) + W);
    FieldIO<F>::store(table + i * ts, x);
    FieldIO<F>::store(table + i * ts + W, y);
    XYZZ<F> acc = XYZZ<F>::from_affine(x, y);
    if (x.is_zero() && y.is_zero()) acc = XYZZ<F>::infinity();  // infinity placeholder: never addressed (scalar zeroed)
    for (unsigned w = 1; w < nwin; w++) {
        for (unsigned k = 0; k < c; k++) acc = XYZZ<F>::dbl(acc);
        F ox = F::zero(), oy = F::zero();
        if (!acc.is_inf()) {
            F inv = F::inv_fermat(F::mul(acc.zz, acc.zzz));
            ox = F::mul(acc.x, F::mul(inv, acc.zzz));
            oy = F::mul(acc.y, F::mul(inv, acc.zz));
            acc = XYZZ<F>::from_affine(ox, oy);
        }
        FieldIO<F>::store(table + ((size_t)w * n + i) * ts, ox);
        FieldIO<F>::store(table + ((size_t)w * n + i) * ts + W, oy);
    }
}
cudaError_t msm_precompute_table(int curve, uint32_t* table, unsigned tstride, const uint32_t* bases, size_t n, unsigned c, unsigned nwin,
                                 cudaStream_t st) {
    if (!n) return cudaSuccess;
    unsigned blocks = (unsigned)((n + 127) / 128);
    if (curve == 1) {
        k_msm_precompute<Fq><<<blocks, 128, 0, st>>>(table, tstride, bases, n, c, nwin);
    } else {
        k_msm_precompute<Fq2><<<blocks, 128, 0, st>>>(table, tstride, bases, n, c, nwin);
    }
    CZK_LAUNCHED();
    return cudaGetLastError();
}

// ------------------------------------------------------------------ synthetic / test input generator
struct U256 {
    uint64_t v[4];
};
constexpr int PROG_CHUNK = 64;
template <class F>
__device__ __forceinline__ XYZZ<F> scalar_mul_affine(const uint32_t* kb, const F& bx, const F& by) {
    XYZZ<F> acc = XYZZ<F>::infinity();
    for (int bit = 252; bit >= 0; bit--) {
        acc = XYZZ<F>::dbl(acc);
        if ((kb[bit >> 5] >> (bit & 31)) & 1) acc.add_affine(bx, by);
    }
    return acc;
}
// out[i] = (k0 + i kstep + i^2 kquad) * base.  A thread computes its chunk's first point and first difference by
// double-and-add, then runs the two finite differences: P_{i+1} = P_i + D_i, D_{i+1} = D_i + 2 kquad base.
template <class F>
__global__ void __launch_bounds__(128) k_gen_progression(uint32_t* __restrict__ out_xy, const uint32_t* __restrict__ base_xy,
                                                          const uint32_t* __restrict__ quad2_xy, U256 k0, U256 kstep, U256 kquad,
                                                          size_t n) {
    constexpr int W = FieldIO<F>::W;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * PROG_CHUNK;
    if (lo >= n) return;
    Fr k0m, ksm, kqm, lom = Fr::zero();
#pragma unroll
    for (int i = 0; i < 4; i++) {
        k0m.l[2 * i] = (uint32_t)k0.v[i];
        k0m.l[2 * i + 1] = (uint32_t)(k0.v[i] >> 32);
        ksm.l[2 * i] = (uint32_t)kstep.v[i];
        ksm.l[2 * i + 1] = (uint32_t)(kstep.v[i] >> 32);
        kqm.l[2 * i] = (uint32_t)kquad.v[i];
        kqm.l[2 * i + 1] = (uint32_t)(kquad.v[i] >> 32);
    }
    lom.l[0] = (uint32_t)lo;
    lom.l[1] = (uint32_t)((uint64_t)lo >> 32);
    k0m = Fr::to_mont(k0m);
    ksm = Fr::to_mont(ksm);
    kqm = Fr::to_mont(kqm);
    lom = Fr::to_mont(lom);
    // k = k0 + lo ks + lo^2 kq ;  dk = ks + (2 lo + 1) kq
    Fr k = Fr::from_mont(Fr::add(k0m, Fr::mul(lom, Fr::add(ksm, Fr::mul(lom, kqm)))));
    Fr dk = Fr::from_mont(Fr::add(ksm, Fr::mul(Fr::add(Fr::dbl(lom), Fr::one()), kqm)));
    uint32_t kb[8], db[8];
#pragma unroll
    for (int i = 0; i < 8; i++) kb[i] = k.l[i], db[i] = dk.l[i];
    F bx = FieldIO<F>::load(base_xy), by = FieldIO<F>::load(base_xy + W);
    F qx = FieldIO<F>::load(quad2_xy), qy = FieldIO<F>::load(quad2_xy + W);
    XYZZ<F> acc = scalar_mul_affine<F>(kb, bx, by);
    XYZZ<F> dif = scalar_mul_affine<F>(db, bx, by);
    size_t hi = lo + PROG_CHUNK < n ? lo + PROG_CHUNK : n;
    for (size_t i = lo; i < hi; i++) {
        F ox, oy;
        if (acc.is_inf()) {
            ox = F::zero();
            oy = F::one();
        } else {
            F inv = F::inv_fermat(F::mul(acc.zz, acc.zzz));
            ox = F::mul(acc.x, F::mul(inv, acc.zzz));
            oy = F::mul(acc.y, F::mul(inv, acc.zz));
        }
        FieldIO<F>::store(out_xy + i * (2 * W), ox);
        FieldIO<F>::store(out_xy + i * (2 * W) + W, oy);
        acc.add(dif);
        dif.add_affine(qx, qy);
    }
}

cudaError_t ec_gen_progression_dev(int curve, uint32_t* out_xy, const uint32_t* base_xy, const uint32_t* quad2_xy,
                                   const uint64_t k0_canon[4], const uint64_t kstep_canon[4], const uint64_t kquad_canon[4],
                                   size_t n, cudaStream_t st) {
    U256 a, b, q;
    for (int i = 0; i < 4; i++) {
        a.v[i] = k0_canon[i];
        b.v[i] = kstep_canon[i];
        q.v[i] = kquad_canon[i];
    }
    size_t threads = (n + PROG_CHUNK - 1) / PROG_CHUNK;
    unsigned blocks = (unsigned)((threads + 127) / 128);
    if (!blocks) return cudaSuccess;
    if (curve == 1) {
        k_gen_progression<Fq><<<blocks, 128, 0, st>>>(out_xy, base_xy, quad2_xy, a, b, q, n);
    } else {
        k_gen_progression<Fq2><<<blocks, 128, 0, st>>>(out_xy, base_xy, quad2_xy, a, b, q, n);
    }
    CZK_LAUNCHED();
    return cudaGetLastError();
}

}  // namespace czk
