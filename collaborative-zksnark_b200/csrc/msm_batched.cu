// Bucket accumulation by batched affine additions (the fast path of step 4 in msm.cuh).
//
// The reference adds every point of a bucket into a Jacobian accumulator (variable_base.rs:66-78, madd-2007-bl,
// 7M + 4S per point); the XYZZ kernel in msm.cu does the same walk at 8M + 2S.  Here a bucket is summed as a binary
// tree of AFFINE additions instead: round r halves every bucket's run (pairs are added, an odd leftover is copied),
// and all the additions of a round - tens of millions, all independent - share inversions through Montgomery's
// trick, so one addition costs 6 field products (1 for the running product, 2 to peel its inverse back off, 3 for
// lambda, lambda^2 and y3) instead of 10.  log2(longest bucket) rounds; the work halves every round.
//
// Layout: round r's points are an array in which bucket b owns the slots [S_r(b), S_r(b) + L_r(b)) with
//     L_r(b) = ceil(L_0(b) / 2^r),   S_{r+1}(b) = floor(S_r(b) / 2) + b,   S_0 / L_0 = the counting sort's runs,
// which never overlap (the + b pays for the rounding) and need no per-round scan: a thread finds its bucket with one
// binary search over the closed form and then walks.  Round 0 reads the bases through the sorted (index | sign) list.
//
// One thread owns B consecutive output slots: it multiplies up the B differences x2 - x1 (prefix products to a
// scratch array), the block multiplies the 128 thread products up a tree in shared memory, ONE thread inverts the
// root (binary extended Euclid, the reference's own inversion algorithm, macros.rs:368-422), the inverses flow back
// down the tree, and each thread peels its B inverses off backwards while it forms the sums.
//
// Affine addition has no formula for P + P or P + (-P) (x2 == x1).  Those cannot occur between distinct CRS points
// and random partial sums, but they do occur in adversarial inputs (a repeated base, P and -P in one bucket): the
// kernel then raises a flag, and the XYZZ kernel - which handles every case the reference does - recomputes the
// buckets.  Both kernels are always enqueued; the one that is not needed returns at once (no host round trip).
#include "msm.cuh"
#include <cstdlib>

#include "fq_inverse.cuh"
#include "launch_count.hpp"
#include "msm_io.cuh"

namespace czk {

struct BatGeom {
    const uint32_t* ends;  // bucket end offsets in the sorted list (after the scatter)
    const uint32_t* hist;  // bucket lengths L_0
    uint32_t nb;           // buckets (all windows)
    int r;                 // the round whose array is the INPUT
    uint32_t bstride;      // words between consecutive point records of `bases` (round 0 only)
};
__device__ __forceinline__ uint32_t bat_start(const BatGeom& g, uint32_t b, int r) {
    uint32_t s = g.ends[b] - g.hist[b];
    for (int i = 0; i < r; i++) s = (s >> 1) + b;
    return s;
}
__device__ __forceinline__ uint32_t bat_len(const BatGeom& g, uint32_t b, int r) {
    return (uint32_t)(((uint64_t)g.hist[b] + ((1ull << r) - 1)) >> r);
}

// How many halving rounds a bucket set needs is decided ON THE DEVICE from the longest bucket (a word the sort leaves in
// the workspace), so the host enqueues a whole MSM without reading anything back: it launches as many round kernels as
// the worst case needs, and a round whose input already has at most `walk` points per bucket returns at once.
__host__ __device__ __forceinline__ uint32_t bat_len_after(uint32_t maxlen, int r) {
    return (uint32_t)(((uint64_t)maxlen + ((1ull << r) - 1)) >> r);
}
__host__ __device__ __forceinline__ bool bat_round_runs(int r, uint32_t maxlen, uint32_t walk) {
    return r == 0 || bat_len_after(maxlen, r) > walk;  // round 0 always runs: it turns (index | sign) entries into points
}
__host__ __device__ __forceinline__ int bat_rounds_needed(uint32_t maxlen, uint32_t walk, int launched) {
    int rounds = 1;
    while (rounds < launched && bat_len_after(maxlen, rounds) > walk) rounds++;
    return rounds;
}

// ------------------------------------------------------------------ inversion (one thread per block)
// a^-1 for a != 0, Montgomery form in and out: the binary extended Euclidean algorithm of the reference
// (algebra/ff/src/fields/macros.rs:368-422) on 32-bit limbs; b starts at R^2 so the result is already in Montgomery form.
__device__ __noinline__ Fq fq_inverse_binary(const Fq& a) {
    constexpr int N = 12;
    uint32_t u[N], v[N], b[N], c[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        u[i] = a.l[i];
        v[i] = FqParams::mod(i);
        b[i] = FqParams::r2(i);
        c[i] = 0;
    }
    auto is_one = [](const uint32_t* x) {
        uint32_t o = x[0] ^ 1u;
#pragma unroll
        for (int i = 1; i < N; i++) o |= x[i];
        return o == 0;
    };
    auto halve = [](uint32_t* x, uint32_t* y) {  // x >>= 1 ; y = y / 2 mod p
#pragma unroll
        for (int i = 0; i < N - 1; i++) x[i] = __funnelshift_r(x[i], x[i + 1], 1);
        x[N - 1] >>= 1;
        if (y[0] & 1u) {  // y + p < 2^378: no carry out of the top limb
            y[0] = add_cc(y[0], FqParams::mod(0));
#pragma unroll
            for (int i = 1; i < N - 1; i++) y[i] = addc_cc(y[i], FqParams::mod(i));
            y[N - 1] = addc(y[N - 1], FqParams::mod(N - 1));
        }
#pragma unroll
        for (int i = 0; i < N - 1; i++) y[i] = __funnelshift_r(y[i], y[i + 1], 1);
        y[N - 1] >>= 1;
    };
    auto sub_pair = [](uint32_t* x, const uint32_t* y, uint32_t* s, const uint32_t* t) {  // x -= y ; s = s - t mod p
        x[0] = sub_cc(x[0], y[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) x[i] = subc_cc(x[i], y[i]);
        x[N - 1] = subc(x[N - 1], y[N - 1]);
        s[0] = sub_cc(s[0], t[0]);
#pragma unroll
        for (int i = 1; i < N; i++) s[i] = subc_cc(s[i], t[i]);
        uint32_t borrow = subc(0, 0);
        s[0] = add_cc(s[0], FqParams::mod(0) & borrow);
#pragma unroll
        for (int i = 1; i < N - 1; i++) s[i] = addc_cc(s[i], FqParams::mod(i) & borrow);
        s[N - 1] = addc(s[N - 1], FqParams::mod(N - 1) & borrow);
    };
    auto less = [](const uint32_t* x, const uint32_t* y) {  // x < y
        uint32_t t = sub_cc(x[0], y[0]);
#pragma unroll
        for (int i = 1; i < N; i++) t = subc_cc(x[i], y[i]);
        (void)t;
        return subc(0, 0) != 0;
    };
#pragma unroll 1
    while (!is_one(u) && !is_one(v)) {
#pragma unroll 1
        while ((u[0] & 1u) == 0) halve(u, b);
#pragma unroll 1
        while ((v[0] & 1u) == 0) halve(v, c);
        if (less(v, u)) sub_pair(u, v, b, c);
        else sub_pair(v, u, c, b);
    }
    Fq r;
    const bool first = is_one(u);
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = first ? b[i] : c[i];
    return r;
}
// BAT_BINGCD=1 selects the batched-step binary GCD of fq_inverse.cuh (same result, ~4x fewer dependent instructions;
// checked on the host by the emulation tests, not yet measured on the device - see DESIGN.md section 7)
#ifndef BAT_BINGCD
#define BAT_BINGCD 0
#endif
__device__ __noinline__ Fq fq_inverse_bingcd(const Fq& a) { return BinGcd<FqParams>::inverse(a); }
__device__ __forceinline__ Fq fq_inverse_one(const Fq& a) { return BAT_BINGCD ? fq_inverse_bingcd(a) : fq_inverse_binary(a); }
__device__ __forceinline__ Fq field_inverse(const Fq& a) { return fq_inverse_one(a); }
// 1 / (c0 + c1 u) = (c0 - c1 u) / (c0^2 + 5 c1^2)   (quadratic_extension.rs:308-324)
__device__ __forceinline__ Fq2 field_inverse(const Fq2& a) {
    Fq norm = Fq::sub(Fq::mul_ni(a.c0, a.c0), Fq2::mul_by_nonresidue(Fq::mul_ni(a.c1, a.c1)));
    Fq ni = fq_inverse_one(norm);
    return Fq2{Fq::mul_ni(a.c0, ni), Fq::neg(Fq::mul_ni(a.c1, ni))};
}
// exported for the parity test of the inversion itself
__global__ void k_fq_inverse(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fq a = FieldIO<Fq>::load(in + i * 12);
    FieldIO<Fq>::store(out + i * 12, a.is_zero() ? a : fq_inverse_one(a));
}
cudaError_t fq_inverse_batch(const uint32_t* in, uint32_t* out, size_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    k_fq_inverse<<<(unsigned)((n + 31) / 32), 32, 0, st>>>(in, out, n); CZK_LAUNCHED();
    return cudaGetLastError();
}

// ------------------------------------------------------------------ shared-memory product tree (limb-major)
template <class F>
struct TreeIO;
template <>
struct TreeIO<Fq> {
    __device__ __forceinline__ static Fq ld(const uint32_t* t, unsigned i) {
        Fq r;
#pragma unroll
        for (int w = 0; w < 12; w++) r.l[w] = t[w * 256 + i];
        return r;
    }
    __device__ __forceinline__ static void st(uint32_t* t, unsigned i, const Fq& v) {
#pragma unroll
        for (int w = 0; w < 12; w++) t[w * 256 + i] = v.l[w];
    }
};
template <>
struct TreeIO<Fq2> {
    __device__ __forceinline__ static Fq2 ld(const uint32_t* t, unsigned i) {
        return Fq2{TreeIO<Fq>::ld(t, i), TreeIO<Fq>::ld(t + 12 * 256, i)};
    }
    __device__ __forceinline__ static void st(uint32_t* t, unsigned i, const Fq2& v) {
        TreeIO<Fq>::st(t, i, v.c0);
        TreeIO<Fq>::st(t + 12 * 256, i, v.c1);
    }
};
// products inside the two hot loops: inlined (for Fq2 the out-of-line bodies cost a round trip through local memory per
// operand; measured below)
template <int INL>
__device__ __forceinline__ Fq hot_mul(const Fq& a, const Fq& b) { return Fq::mul(a, b); }
template <int INL>
__device__ __forceinline__ Fq hot_sqr(const Fq& a) { return Fq::sqr(a); }
template <int INL>
__device__ __forceinline__ Fq2 hot_mul(const Fq2& a, const Fq2& b) { return INL ? Fq2::mul_inl(a, b) : Fq2::mul(a, b); }
template <int INL>
__device__ __forceinline__ Fq2 hot_sqr(const Fq2& a) { return INL ? Fq2::sqr_inl(a) : Fq2::sqr(a); }
// products outside the two hot loops go through out-of-line bodies (code size)
__device__ __forceinline__ Fq cold_mul(const Fq& a, const Fq& b) { return Fq::mul_ni(a, b); }
__device__ __forceinline__ Fq2 cold_mul(const Fq2& a, const Fq2& b) { return Fq2::mul(a, b); }

// ------------------------------------------------------------------ walking the slot geometry
struct BatCursor {
    uint32_t b, s_in, len_in, s_out, next_s_out;
    __device__ __forceinline__ uint32_t len_out() const { return (len_in + 1) >> 1; }
};
__device__ __forceinline__ void bat_seek(const BatGeom& g, BatCursor& c, uint32_t b) {
    c.b = b;
    c.s_in = bat_start(g, b, g.r);
    c.len_in = bat_len(g, b, g.r);
    c.s_out = (c.s_in >> 1) + b;
    c.next_s_out = b + 1 < g.nb ? (bat_start(g, b + 1, g.r) >> 1) + b + 1 : 0xffffffffu;
}

constexpr int BAT_THREADS = 128;
#ifndef BAT_INLINE_FQ2
#define BAT_INLINE_FQ2 1
#endif
// ------------------------------------------------------------------ one round
// How the operands reach a thread was tuned against the ncu source view (profiles/r1_summary.md).  With 128 registers
// per thread (4 blocks of 128 per SM) nothing can be prefetched into registers: a first version that fetched the sorted
// entries one slot ahead had them spilled at once, which waits for the load just the same, and a third of all stall
// samples sat on (i) those entries, (ii) the prefix product pre[j-1], loaded at its point of use, (iii) the y negation
// placed right behind the gathers.  So the SMALL operands of the next slot - its two sorted entries and its prefix
// product - travel by cp.async into per-thread shared-memory cells (double buffered by slot parity: a copy never
// targets a cell that is being read) while the current slot is computed; no register is held while they fly.  The
// 96-byte point gathers go straight to registers, issued at the top of a slot and first consumed after the 1/d
// product.  (Staging the gathers through shared memory as well was 18 % slower: 30 more LDGSTS/LDS per slot and the
// copy queue throttles.)
__device__ __forceinline__ void cp_async16(uint4* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void words_to(Fq& r, const uint32_t* w) {
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = w[i];
}
__device__ __forceinline__ void words_to(Fq2& r, const uint32_t* w) {
    words_to(r.c0, w);
    words_to(r.c1, w + 12);
}
// one field element per thread and parity: chunk c of thread tid at cell[(par * CH + c) * BAT_THREADS + tid]
template <class F, int THREADS = BAT_THREADS>
struct StageIO {
    static constexpr int W = FieldIO<F>::W, CH = W / 4;
    __device__ __forceinline__ static void fetch(uint4* cell, int par, unsigned tid, const uint32_t* g) {
#pragma unroll
        for (int c = 0; c < CH; c++) cp_async16(cell + (par * CH + c) * THREADS + tid, g + 4 * c);
    }
    __device__ __forceinline__ static F ld(const uint4* cell, int par, unsigned tid) {
        uint32_t w[W];
#pragma unroll
        for (int c = 0; c < CH; c++) {
            const uint4 v = cell[(par * CH + c) * THREADS + tid];
            w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
        }
        F r;
        words_to(r, w);
        return r;
    }
};
template <class F>
constexpr size_t bat_smem_bytes() {  // product tree + 2 prefix cells + 2 x 2 entry cells per thread
    return (size_t)FieldIO<F>::W * 256 * 4 + (size_t)2 * (FieldIO<F>::W / 4) * BAT_THREADS * 16 + (size_t)4 * BAT_THREADS * 4;
}

template <class F, bool FIRST, int MINB>
__global__ void __launch_bounds__(BAT_THREADS, MINB)
    k_bat_round(BatGeom g, const uint32_t* __restrict__ bases, const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ in,
                uint32_t* __restrict__ out, uint32_t* __restrict__ prefix, uint32_t* __restrict__ flag, const int B,
                const uint32_t* __restrict__ maxlen_p, const uint32_t walk) {
    constexpr int W = FieldIO<F>::W, CH = W / 4;
    if (!bat_round_runs(g.r, *maxlen_p, walk)) return;  // the halving already brought every bucket under `walk` points
    extern __shared__ uint4 bat_smem[];
    uint32_t* const tree = reinterpret_cast<uint32_t*>(bat_smem);
    uint4* const pcell = bat_smem + (W * 256) / 4;
    uint32_t* const ecell = reinterpret_cast<uint32_t*>(pcell + 2 * CH * BAT_THREADS);  // [par][which][tid]
    const unsigned tid = threadIdx.x;
    const uint32_t used = (bat_start(g, g.nb - 1, g.r) >> 1) + (g.nb - 1) + ((bat_len(g, g.nb - 1, g.r) + 1) >> 1);
    const size_t block_first = (size_t)blockIdx.x * BAT_THREADS * B;
    if (block_first >= used) return;
    const size_t o0 = block_first + (size_t)tid * B;
    uint32_t* const pre = prefix + ((size_t)blockIdx.x * B * BAT_THREADS + tid) * W;
    constexpr size_t PRE_STRIDE = (size_t)BAT_THREADS * W;

    BatCursor c;
    {
        uint32_t lo = 0, hi = g.nb - 1;
        const uint32_t target = o0 > 0xfffffffeull ? 0xfffffffeu : (uint32_t)o0;
        while (lo < hi) {
            uint32_t mid = lo + (hi - lo + 1) / 2;
            uint32_t s = (bat_start(g, mid, g.r) >> 1) + mid;
            if (s <= target) lo = mid;
            else hi = mid - 1;
        }
        bat_seek(g, c, lo);
    }
    const F one = F::one();
    // descriptor of the slot whose small operands are in flight: bit 0 live, bit 1 pair; m_pos = its first input position
    uint32_t m_fl = 0, m_pos = 0;
    auto point_ptr = [&](uint32_t pos, uint32_t e) -> const uint32_t* {
        return FIRST ? bases + (size_t)(e & 0x7fffffffu) * g.bstride : in + (size_t)pos * (2 * W);
    };
    auto ldp = [&](const uint32_t* q) -> F { return FIRST ? FieldIO<F>::load(q) : FieldIO<F>::load_rw(q); };
    auto fetch_entries = [&](int par) {
        if (FIRST && (m_fl & 1u)) cp_async4(ecell + (par * 2 + 0) * BAT_THREADS + tid, sorted + m_pos);
        if (FIRST && (m_fl & 2u)) cp_async4(ecell + (par * 2 + 1) * BAT_THREADS + tid, sorted + m_pos + 1);
    };
    auto entry = [&](int par, int which) -> uint32_t { return FIRST ? ecell[(par * 2 + which) * BAT_THREADS + tid] : 0u; };
    // ---- phase 1: running product of the differences x2 - x1
    auto meta1 = [&](size_t o) {
        while (o >= c.next_s_out) bat_seek(g, c, c.b + 1);
        const uint32_t k = (uint32_t)(o - c.s_out);
        const bool pair = k < c.len_out() && 2 * k + 1 < c.len_in;
        m_pos = c.s_in + 2 * k;
        m_fl = pair ? 3u : 0u;
    };
    F acc = one;
    F nx1, nx2;
    bool npair;
    {
        meta1(o0);
        npair = m_fl & 2u;
        if (npair) {
            const uint32_t e1 = FIRST ? sorted[m_pos] : 0u, e2 = FIRST ? sorted[m_pos + 1] : 0u;
            nx1 = ldp(point_ptr(m_pos, e1));
            nx2 = ldp(point_ptr(m_pos + 1, e2));
        }
        if (B > 1) {
            meta1(o0 + 1);
            fetch_entries(1);
        }
        cp_async_commit();
    }
#pragma unroll 1
    for (int j = 0; j < B; j++) {
        const bool pair = npair;
        F x1 = nx1, x2 = nx2;
        if (j + 1 < B) {
            const int par = (j + 1) & 1;
            cp_async_wait_all();
            npair = m_fl & 2u;
            if (npair) {
                nx1 = ldp(point_ptr(m_pos, entry(par, 0)));
                nx2 = ldp(point_ptr(m_pos + 1, entry(par, 1)));
            }
            if (j + 2 < B) {
                meta1(o0 + j + 2);
                fetch_entries(par ^ 1);
            }
            cp_async_commit();
        }
        F d = one;
        if (pair) {
            d = F::sub(x2, x1);
            if (d.is_zero()) {
                atomicOr(flag, 1u);
                d = one;
            }
        }
        acc = j == 0 ? d : hot_mul<BAT_INLINE_FQ2>(acc, d);
        FieldIO<F>::store(pre + (size_t)j * PRE_STRIDE, acc);
    }
    // ---- phase 2: one inversion for the block (the last slot's small operands fly meanwhile)
    TreeIO<F>::st(tree, 128 + tid, acc);
    __syncthreads();
    auto meta3 = [&](size_t o) {
        while (o < c.s_out) bat_seek(g, c, c.b - 1);
        const uint32_t k = (uint32_t)(o - c.s_out);
        const bool live = k < c.len_out(), pair = live && 2 * k + 1 < c.len_in;
        m_pos = c.s_in + 2 * k;
        m_fl = (live ? 1u : 0u) | (pair ? 2u : 0u);
    };
    auto fetch_small = [&](int j) {  // slot j's entries and the prefix product below it, into the cells of j's parity
        fetch_entries(j & 1);
        if (j > 0) StageIO<F>::fetch(pcell, j & 1, tid, pre + (size_t)(j - 1) * PRE_STRIDE);
        cp_async_commit();
    };
    meta3(o0 + B - 1);
    fetch_small(B - 1);
#pragma unroll 1
    for (unsigned width = 64; width >= 1; width >>= 1) {
        if (tid < width) {
            unsigned i = width + tid;
            TreeIO<F>::st(tree, i, cold_mul(TreeIO<F>::ld(tree, 2 * i), TreeIO<F>::ld(tree, 2 * i + 1)));
        }
        __syncthreads();
    }
    if (tid == 0) TreeIO<F>::st(tree, 1, field_inverse(TreeIO<F>::ld(tree, 1)));
    __syncthreads();
#pragma unroll 1
    for (unsigned width = 1; width <= 64; width <<= 1) {
        if (tid < width) {
            unsigned i = width + tid;
            F inv_i = TreeIO<F>::ld(tree, i), l = TreeIO<F>::ld(tree, 2 * i), r = TreeIO<F>::ld(tree, 2 * i + 1);
            TreeIO<F>::st(tree, 2 * i, cold_mul(inv_i, r));
            TreeIO<F>::st(tree, 2 * i + 1, cold_mul(inv_i, l));
        }
        __syncthreads();
    }
    F inv = TreeIO<F>::ld(tree, 128 + tid);
    // ---- phase 3: peel the inverses off backwards and form the sums
#pragma unroll 1
    for (int j = B - 1; j >= 0; j--) {
        const size_t o = o0 + j;
        const bool live = m_fl & 1u, pair = m_fl & 2u;
        const int par = j & 1;
        cp_async_wait_all();
        const uint32_t e1 = entry(par, 0), e2 = entry(par, 1);
        F x1, y1, x2, y2, d = one;
        if (live) {
            const uint32_t* p = point_ptr(m_pos, e1);
            x1 = ldp(p);
            y1 = ldp(p + W);
        }
        if (pair) {
            const uint32_t* p = point_ptr(m_pos + 1, e2);
            x2 = ldp(p);
            y2 = ldp(p + W);
        }
        F dinv = inv;
        if (j > 0) {
            const F pj = StageIO<F>::ld(pcell, par, tid);
            meta3(o - 1);
            fetch_small(j - 1);
            dinv = hot_mul<BAT_INLINE_FQ2>(inv, pj);  // runs under the gathers' latency
        }
        // order chosen for register pressure: lambda first (y2 and 1/d die), then d and the running inverse
        if (FIRST && live && (e1 >> 31)) y1 = F::neg(y1);
        F lam;
        if (pair) {
            if (FIRST && (e2 >> 31)) y2 = F::neg(y2);
            lam = hot_mul<BAT_INLINE_FQ2>(F::sub(y2, y1), dinv);
            d = F::sub(x2, x1);
            if (d.is_zero()) d = one;
        }
        if (j > 0) inv = hot_mul<BAT_INLINE_FQ2>(inv, d);
        uint32_t* dst = out + o * (2 * W);
        if (pair) {
            F x3 = F::sub(F::sub(hot_sqr<BAT_INLINE_FQ2>(lam), x1), x2);
            F y3 = F::sub(hot_mul<BAT_INLINE_FQ2>(lam, F::sub(x1, x3)), y1);
            FieldIO<F>::store(dst, x3);
            FieldIO<F>::store(dst + W, y3);
        } else if (live) {
            FieldIO<F>::store(dst, x1);
            FieldIO<F>::store(dst + W, y1);
        }
    }
}

// ------------------------------------------------------------------ one round on G2: two lanes per slot
// The same round for Fq2 points with every slot owned by a PAIR of adjacent lanes: the even lane holds the c0 component
// of every Fq2 value, the odd lane c1.  Sums, differences, negations and zero tests of Fq2 are component-wise, so they
// split for free; a product needs both components of both operands, which the partner supplies by 24 warp shuffles,
// and each lane then forms ONE lazily reduced sum of two base-field products (Fp::mul_sum2, one Montgomery reduction):
//     c0 = a0 b0 + (-5 a1) b1          c1 = a1 b0 + a0 b1          (quadratic_extension.rs:569-583 without Karatsuba)
// 2 x (2 * 144 + 132) = 840 wide multiply-adds per Fq2 product against 3 x 276 = 828 for Karatsuba on one lane - the
// same arithmetic - but a lane carries HALF the state: 128 registers instead of 255, so 4 warps per scheduler are
// resident instead of 2 (the one-lane kernel sat at 41 % of the base-field product rate, profiles/r1_summary.md).
// Shuffles need the whole warp, so the products of a slot are executed unconditionally (a slot that only copies a
// leftover point computes on zeros and stores the copy); every condition that guards a store is identical in both
// lanes of a pair.
constexpr int G2L_THREADS = 2 * BAT_THREADS;
__device__ __forceinline__ Fq g2l_mul(const Fq& a, const Fq& b, const unsigned h) {
    uint32_t pa[12];
    {
        uint32_t sa[12];
        Fq::neg_times5_unreduced(a, sa);  // the odd lane hands out -5 a1 (as 5 (p - a1) < 2^380), the even lane a0
#pragma unroll
        for (int i = 0; i < 12; i++) pa[i] = __shfl_xor_sync(0xffffffffu, h ? sa[i] : a.l[i], 1);
    }
    // multiplier limbs row by row: b0 (first product) and b1 (second product) in BOTH lanes, one shuffle per row
    auto rows = [&](int i, uint32_t& b0, uint32_t& b1) {
        const uint32_t pb = __shfl_xor_sync(0xffffffffu, b.l[i], 1);
        b0 = h ? pb : b.l[i];
        b1 = h ? b.l[i] : pb;
    };
    return Fq::mul_sum2_f(a, pa, rows);
}
__device__ __forceinline__ bool g2l_both(bool mine) {  // true iff the condition holds in both lanes of the pair
    const unsigned theirs = __shfl_xor_sync(0xffffffffu, mine ? 1u : 0u, 1);  // never behind a short-circuit: every lane shuffles
    return mine && theirs != 0u;
}
constexpr size_t g2l_smem_bytes() {  // Fq2 product tree + 2 prefix cells (one component) + 2 x 2 entry cells, per thread
    return (size_t)24 * 256 * 4 + (size_t)2 * 3 * G2L_THREADS * 16 + (size_t)4 * G2L_THREADS * 4;
}

template <bool FIRST, int MINB>
__global__ void __launch_bounds__(G2L_THREADS, MINB)
    k_bat_round_g2l(BatGeom g, const uint32_t* __restrict__ bases, const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ in,
                    uint32_t* __restrict__ out, uint32_t* __restrict__ prefix, uint32_t* __restrict__ flag, const int B,
                    const uint32_t* __restrict__ maxlen_p, const uint32_t walk) {
    constexpr int W = 24, HW = 12, CH = 3;  // words of an Fq2 element, of one component, 16-byte chunks of a component
    if (!bat_round_runs(g.r, *maxlen_p, walk)) return;
    typedef StageIO<Fq, G2L_THREADS> Stage;
    extern __shared__ uint4 bat_smem[];
    uint32_t* const tree = reinterpret_cast<uint32_t*>(bat_smem);
    uint4* const pcell = bat_smem + (W * 256) / 4;
    uint32_t* const ecell = reinterpret_cast<uint32_t*>(pcell + 2 * CH * G2L_THREADS);  // [par][which][tid]
    const unsigned tid = threadIdx.x, ow = tid >> 1, h = tid & 1;
    const uint32_t used = (bat_start(g, g.nb - 1, g.r) >> 1) + (g.nb - 1) + ((bat_len(g, g.nb - 1, g.r) + 1) >> 1);
    const size_t block_first = (size_t)blockIdx.x * BAT_THREADS * B;
    if (block_first >= used) return;
    const size_t o0 = block_first + (size_t)ow * B;
    uint32_t* const pre = prefix + ((size_t)blockIdx.x * B * BAT_THREADS + ow) * W + HW * h;
    constexpr size_t PRE_STRIDE = (size_t)BAT_THREADS * W;
    uint32_t* const mytree = tree + HW * 256 * h;  // this lane's component of the limb-major Fq2 tree

    BatCursor c;
    {
        uint32_t lo = 0, hi = g.nb - 1;
        const uint32_t target = o0 > 0xfffffffeull ? 0xfffffffeu : (uint32_t)o0;
        while (lo < hi) {
            uint32_t mid = lo + (hi - lo + 1) / 2;
            uint32_t s = (bat_start(g, mid, g.r) >> 1) + mid;
            if (s <= target) lo = mid;
            else hi = mid - 1;
        }
        bat_seek(g, c, lo);
    }
    const Fq one = h ? Fq::zero() : Fq::one();  // this lane's component of 1
    uint32_t m_fl = 0, m_pos = 0;
    auto point_ptr = [&](uint32_t pos, uint32_t e) -> const uint32_t* {
        return (FIRST ? bases + (size_t)(e & 0x7fffffffu) * g.bstride : in + (size_t)pos * (2 * W)) + HW * h;
    };
    auto ldp = [&](const uint32_t* q) -> Fq { return FIRST ? FieldIO<Fq>::load(q) : FieldIO<Fq>::load_rw(q); };
    auto fetch_entries = [&](int par) {
        if (FIRST && (m_fl & 1u)) cp_async4(ecell + (par * 2 + 0) * G2L_THREADS + tid, sorted + m_pos);
        if (FIRST && (m_fl & 2u)) cp_async4(ecell + (par * 2 + 1) * G2L_THREADS + tid, sorted + m_pos + 1);
    };
    auto entry = [&](int par, int which) -> uint32_t { return FIRST ? ecell[(par * 2 + which) * G2L_THREADS + tid] : 0u; };
    // ---- phase 1: running product of the differences x2 - x1
    auto meta1 = [&](size_t o) {
        while (o >= c.next_s_out) bat_seek(g, c, c.b + 1);
        const uint32_t k = (uint32_t)(o - c.s_out);
        const bool pair = k < c.len_out() && 2 * k + 1 < c.len_in;
        m_pos = c.s_in + 2 * k;
        m_fl = pair ? 3u : 0u;
    };
    Fq acc = one;
    Fq nx1 = Fq::zero(), nx2 = Fq::zero();
    bool npair;
    {
        meta1(o0);
        npair = m_fl & 2u;
        if (npair) {
            const uint32_t e1 = FIRST ? sorted[m_pos] : 0u, e2 = FIRST ? sorted[m_pos + 1] : 0u;
            nx1 = ldp(point_ptr(m_pos, e1));
            nx2 = ldp(point_ptr(m_pos + 1, e2));
        }
        if (B > 1) {
            meta1(o0 + 1);
            fetch_entries(1);
        }
        cp_async_commit();
    }
#pragma unroll 1
    for (int j = 0; j < B; j++) {
        const bool pair = npair;
        const Fq x1 = nx1, x2 = nx2;
        if (j + 1 < B) {
            const int par = (j + 1) & 1;
            cp_async_wait_all();
            npair = m_fl & 2u;
            if (npair) {
                nx1 = ldp(point_ptr(m_pos, entry(par, 0)));
                nx2 = ldp(point_ptr(m_pos + 1, entry(par, 1)));
            }
            if (j + 2 < B) {
                meta1(o0 + j + 2);
                fetch_entries(par ^ 1);
            }
            cp_async_commit();
        }
        Fq d = Fq::sub(x2, x1);
        const bool dz = g2l_both(d.is_zero());
        if (pair && dz) atomicOr(flag, 1u);
        if (!pair || dz) d = one;
        acc = j == 0 ? d : g2l_mul(acc, d, h);
        FieldIO<Fq>::store(pre + (size_t)j * PRE_STRIDE, acc);
    }
    // ---- phase 2: one inversion for the block (the Fq2 tree is walked by single lanes on whole elements: cold code)
    TreeIO<Fq>::st(mytree, 128 + ow, acc);
    __syncthreads();
    auto meta3 = [&](size_t o) {
        while (o < c.s_out) bat_seek(g, c, c.b - 1);
        const uint32_t k = (uint32_t)(o - c.s_out);
        const bool live = k < c.len_out(), pair = live && 2 * k + 1 < c.len_in;
        m_pos = c.s_in + 2 * k;
        m_fl = (live ? 1u : 0u) | (pair ? 2u : 0u);
    };
    auto fetch_small = [&](int j) {
        fetch_entries(j & 1);
        if (j > 0) Stage::fetch(pcell, j & 1, tid, pre + (size_t)(j - 1) * PRE_STRIDE);
        cp_async_commit();
    };
    meta3(o0 + B - 1);
    fetch_small(B - 1);
#pragma unroll 1
    for (unsigned width = 64; width >= 1; width >>= 1) {
        if (tid < width) {
            unsigned i = width + tid;
            TreeIO<Fq2>::st(tree, i, cold_mul(TreeIO<Fq2>::ld(tree, 2 * i), TreeIO<Fq2>::ld(tree, 2 * i + 1)));
        }
        __syncthreads();
    }
    if (tid == 0) TreeIO<Fq2>::st(tree, 1, field_inverse(TreeIO<Fq2>::ld(tree, 1)));
    __syncthreads();
#pragma unroll 1
    for (unsigned width = 1; width <= 64; width <<= 1) {
        if (tid < width) {
            unsigned i = width + tid;
            Fq2 inv_i = TreeIO<Fq2>::ld(tree, i), l = TreeIO<Fq2>::ld(tree, 2 * i), r = TreeIO<Fq2>::ld(tree, 2 * i + 1);
            TreeIO<Fq2>::st(tree, 2 * i, cold_mul(inv_i, r));
            TreeIO<Fq2>::st(tree, 2 * i + 1, cold_mul(inv_i, l));
        }
        __syncthreads();
    }
    Fq inv = TreeIO<Fq>::ld(mytree, 128 + ow);
    // ---- phase 3: peel the inverses off backwards and form the sums
#pragma unroll 1
    for (int j = B - 1; j >= 0; j--) {
        const size_t o = o0 + j;
        const bool live = m_fl & 1u, pair = m_fl & 2u;
        const int par = j & 1;
        cp_async_wait_all();
        const uint32_t e1 = entry(par, 0), e2 = entry(par, 1);
        Fq x1 = Fq::zero(), y1 = Fq::zero(), x2 = Fq::zero(), y2 = Fq::zero();
        if (live) {
            const uint32_t* p = point_ptr(m_pos, e1);
            x1 = ldp(p);
            y1 = ldp(p + W);
        }
        if (pair) {
            const uint32_t* p = point_ptr(m_pos + 1, e2);
            x2 = ldp(p);
            y2 = ldp(p + W);
        }
        Fq dinv = inv;
        if (j > 0) {
            const Fq pj = Stage::ld(pcell, par, tid);
            meta3(o - 1);
            fetch_small(j - 1);
            dinv = g2l_mul(inv, pj, h);  // runs under the gathers' latency
        }
        if (FIRST && live && (e1 >> 31)) y1 = Fq::neg(y1);
        if (FIRST && pair && (e2 >> 31)) y2 = Fq::neg(y2);
        const Fq lam = g2l_mul(Fq::sub(y2, y1), dinv, h);
        Fq d = Fq::sub(x2, x1);
        const bool dz = g2l_both(d.is_zero());  // evaluated by every lane (a shuffle), whatever `pair` says
        if (!pair || dz) d = one;
        if (j > 0) inv = g2l_mul(inv, d, h);
        uint32_t* dst = out + o * (2 * W) + HW * h;
        const Fq x3 = Fq::sub(Fq::sub(g2l_mul(lam, lam, h), x1), x2);
        const Fq y3 = Fq::sub(g2l_mul(lam, Fq::sub(x1, x3), h), y1);
        if (pair) {
            FieldIO<Fq>::store(dst, x3);
            FieldIO<Fq>::store(dst + W, y3);
        } else if (live) {
            FieldIO<Fq>::store(dst, x1);
            FieldIO<Fq>::store(dst + W, y1);
        }
    }
}

// buckets[b] = sum of the (few) points left in bucket b after the halving rounds, as XYZZ for the reduction kernels.
// The late rounds of the tree hold little work but each still costs a block-wide inversion; once every bucket is
// down to BAT_WALK points or fewer, a plain mixed-addition walk (one thread per bucket, every special case of the
// reference's add_assign_mixed handled by XYZZ::add_affine) finishes them in one launch.
constexpr uint32_t BAT_WALK = 24;
template <class F>
__global__ void __launch_bounds__(128) k_bat_finish(BatGeom g, const uint32_t* __restrict__ in_a, const uint32_t* __restrict__ in_b,
                                                     uint32_t* __restrict__ buckets, const uint32_t* __restrict__ flag,
                                                     const uint32_t* __restrict__ maxlen_p, const uint32_t walk, const int launched) {
    constexpr int W = FieldIO<F>::W;
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= g.nb || *flag) return;
    const int rounds = bat_rounds_needed(*maxlen_p, walk, launched);  // round r wrote buffer r & 1
    const uint32_t* in = ((rounds - 1) & 1) ? in_b : in_a;
    XYZZ<F> p = XYZZ<F>::infinity();
    const uint32_t len = bat_len(g, b, rounds);
    const uint32_t* src = in + (size_t)bat_start(g, b, rounds) * (2 * W);
    for (uint32_t i = 0; i < len; i++, src += 2 * W) p.add_affine(FieldIO<F>::load_rw(src), FieldIO<F>::load_rw(src + W));
    store_point<F>(buckets + (size_t)b * (4 * W), p);
}

// slots of the array produced by round r (an upper bound the host can compute): E / 2^(r+1) + 2 nb
static size_t bat_bound(size_t entries, size_t nb, int r_out) { return (entries >> r_out) + 2 * nb + 2; }

// slots per thread: 32 while a round has enough slots to fill the machine, fewer in the short late rounds so that
// their serial per-thread chain (and with it the latency floor of a round) shrinks
constexpr int BAT_B_MAX = 128, BAT_B_MIN = 4;
static int bat_env(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
template <class F>
struct BatTuning;
template <>
struct BatTuning<Fq> {
    static constexpr int MINB = 4;  // 128 registers: 4 blocks of 128 threads per SM (measured 5-8 % faster than 3 at 144)
};
template <>
struct BatTuning<Fq2> {
    static constexpr int MINB = 2;  // 255 registers; 3 blocks at 168 registers spill 450 bytes and measured 12 % slower
};
// Slots per thread for a round with `used` slots.  A block's life is B slot-times plus the block-wide inversion (phase 2,
// ~BETA slot-times: 13 % of a block at B = 64 in the ncu source view), and the grid runs in waves of `resident` blocks, so
// the round costs about ceil(blocks / resident) * (B + BETA): pick the B that makes k waves exactly full for the cheapest
// k.  (Measured on 2^21 G1 terms, round 0 = 15.8 M slots on 592 resident blocks: B = 64 -> 3.3 waves 11.2 ms,
// 96 -> 2.2 waves 11.0 ms, 80 -> 2.6 waves 10.2 ms, 128 -> 1.6 waves 10.3 ms.)
template <class F>
struct BatBeta;
template <>
struct BatBeta<Fq> {
    static constexpr int VALUE = 10;
};
template <>
struct BatBeta<Fq2> {
    static constexpr int VALUE = 4;
};
static int bat_pick_b(size_t used, size_t resident, int beta_dflt) {
    static const int bmax_env = bat_env("CZK_BAT_BMAX", BAT_B_MAX);
    static const int beta_env = bat_env("CZK_BAT_BETA", 0);
    const size_t bmax = bmax_env < BAT_B_MIN ? BAT_B_MIN : (bmax_env > BAT_B_MAX ? BAT_B_MAX : bmax_env);
    const size_t beta = beta_env > 0 ? beta_env : beta_dflt;
    const size_t per_wave = (size_t)BAT_THREADS * resident;
    size_t k0 = (used + per_wave * bmax - 1) / (per_wave * bmax);  // fewest waves that B <= bmax allows
    if (k0 == 0) k0 = 1;
    size_t best_b = bmax, best_cost = ~(size_t)0;
    for (size_t k = k0; k < k0 + 6; k++) {
        size_t b = (used + per_wave * k - 1) / (per_wave * k);
        if (b < BAT_B_MIN) b = BAT_B_MIN;
        if (b > bmax) b = bmax;
        const size_t cost = k * (b + beta);
        if (cost < best_cost) best_cost = cost, best_b = b;
        if (b == BAT_B_MIN) break;
    }
    return (int)best_b;
}
size_t msm_batched_bytes(int curve, size_t entries, size_t nb, size_t* pa, size_t* pb, size_t* pre) {
    const size_t pt = (curve == 1 ? 24 : 48) * 4, el = pt / 2;
    *pa = bat_bound(entries, nb, 1) * pt;
    *pb = bat_bound(entries, nb, 2) * pt;
    size_t slots = bat_bound(entries, nb, 1), per_block = (size_t)BAT_THREADS * BAT_B_MAX;
    *pre = ((slots + per_block - 1) / per_block + 1) * per_block * el;  // a round's last block may overhang by < one block
    return *pa + *pb + *pre;
}

#ifndef BAT_G2_LANES
#define BAT_G2_LANES 1
#endif
constexpr int G2L_MINB = 2;  // 2 blocks of 256 threads at 128 registers
template <class F>
static cudaError_t msm_batched_t(const uint32_t* bases, unsigned bstride, const uint32_t* sorted, const uint32_t* ends, const uint32_t* hist,
                                 size_t nb, size_t entries, const uint32_t* maxlen_p, uint32_t* pa, uint32_t* pb, uint32_t* prefix,
                                 uint32_t* buckets, uint32_t* flag, int sm_count, cudaStream_t st) {
    constexpr int MINB = BatTuning<F>::MINB;
    constexpr bool LANES = BAT_G2_LANES && FieldIO<F>::W == 24;
    // points per bucket at which the tree stops and k_bat_finish walks (measured, 2^20 G2 terms: 24 -> 14.5 ms, 12 -> 14.2, 6 -> 14.2;
    // 2^21 G1 terms: 24 -> 9.63, 12 -> 9.67, 6 -> 9.76)
    static const uint32_t walk = (uint32_t)bat_env(FieldIO<F>::W == 24 ? "CZK_BAT_WALK_G2" : "CZK_BAT_WALK_G1", FieldIO<F>::W == 24 ? 12 : (int)BAT_WALK);
    // rounds to LAUNCH: what the worst case needs (every entry in one bucket), or the usual case plus a margin when the
    // caller caps it; rounds that are not needed return at once (bat_round_runs), and k_bat_finish walks whatever is left
    int launched = 1;
    while (launched < 26 && bat_len_after((uint32_t)(entries > 0xffffffffull ? 0xffffffffu : entries), launched) > walk) launched++;
    BatGeom g{ends, hist, (uint32_t)nb, 0, bstride ? bstride : 2u * FieldIO<F>::W};
    uint32_t* bufs[2] = {pa, pb};
    constexpr size_t smem = LANES ? g2l_smem_bytes() : bat_smem_bytes<F>();
    static const cudaError_t attr = [] {
        cudaError_t e1, e2;
        if (LANES) {
            e1 = cudaFuncSetAttribute(k_bat_round_g2l<true, G2L_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            e2 = cudaFuncSetAttribute(k_bat_round_g2l<false, G2L_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        } else {
            e1 = cudaFuncSetAttribute(k_bat_round<F, true, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            e2 = cudaFuncSetAttribute(k_bat_round<F, false, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        return e1 != cudaSuccess ? e1 : e2;
    }();
    if (attr != cudaSuccess) return attr;
    for (int r = 0; r < launched; r++) {
        g.r = r;
        // the grid and the slots per thread come from the host's bound on the round's slots (entries / 2^(r+1) + 2 nb: within
        // a fraction of a percent of the real count for the rounds that matter); blocks past the real end return at once
        const size_t slots = bat_bound(entries, nb, r + 1);
        const int B = bat_pick_b(slots, (size_t)sm_count * (LANES ? G2L_MINB : MINB), BatBeta<F>::VALUE);
        const size_t per_block = (size_t)BAT_THREADS * B;
        unsigned blocks = (unsigned)((slots + per_block - 1) / per_block);
        uint32_t* dst = bufs[r & 1];
        const uint32_t* src = r ? bufs[(r - 1) & 1] : nullptr;
        if (LANES) {
            if (r == 0) k_bat_round_g2l<true, G2L_MINB><<<blocks, G2L_THREADS, smem, st>>>(g, bases, sorted, nullptr, dst, prefix, flag, B, maxlen_p, walk);
            else k_bat_round_g2l<false, G2L_MINB><<<blocks, G2L_THREADS, smem, st>>>(g, nullptr, nullptr, src, dst, prefix, flag, B, maxlen_p, walk);
        } else if (r == 0) k_bat_round<F, true, MINB><<<blocks, BAT_THREADS, smem, st>>>(g, bases, sorted, nullptr, dst, prefix, flag, B, maxlen_p, walk);
        else k_bat_round<F, false, MINB><<<blocks, BAT_THREADS, smem, st>>>(g, nullptr, nullptr, src, dst, prefix, flag, B, maxlen_p, walk);
        CZK_LAUNCHED();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    k_bat_finish<F><<<(unsigned)((nb + 127) / 128), 128, 0, st>>>(g, pa, pb, buckets, flag, maxlen_p, walk, launched); CZK_LAUNCHED();
    return cudaGetLastError();
}

cudaError_t msm_batched_accumulate(int curve, const uint32_t* bases, unsigned bstride, const uint32_t* sorted, const uint32_t* ends,
                                   const uint32_t* hist, size_t nb, size_t entries, const uint32_t* maxlen_dev, uint32_t* pa,
                                   uint32_t* pb, uint32_t* prefix, uint32_t* buckets, uint32_t* flag, int sm_count, cudaStream_t st) {
    if (curve == 1)
        return msm_batched_t<Fq>(bases, bstride, sorted, ends, hist, nb, entries, maxlen_dev, pa, pb, prefix, buckets, flag, sm_count, st);
    return msm_batched_t<Fq2>(bases, bstride, sorted, ends, hist, nb, entries, maxlen_dev, pa, pb, prefix, buckets, flag, sm_count, st);
}

}  // namespace czk
