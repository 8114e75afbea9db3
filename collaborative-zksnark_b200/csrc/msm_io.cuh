// 16-byte vector loads / stores of field elements and points (shared by the MSM translation units).
#pragma once
#include "ec.cuh"

namespace czk {

// ------------------------------------------------------------------ element I/O (16-byte vector accesses)
template <class F>
struct FieldIO;
template <>
struct FieldIO<Fq> {
    static constexpr int W = 12;
    __device__ __forceinline__ static Fq load(const uint32_t* p) {
        const uint4* q = reinterpret_cast<const uint4*>(p);
        uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
        Fq r;
        r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
        r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
        r.l[8] = c.x; r.l[9] = c.y; r.l[10] = c.z; r.l[11] = c.w;
        return r;
    }
    __device__ __forceinline__ static Fq load_rw(const uint32_t* p) {
        const uint4* q = reinterpret_cast<const uint4*>(p);
        uint4 a = q[0], b = q[1], c = q[2];
        Fq r;
        r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
        r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
        r.l[8] = c.x; r.l[9] = c.y; r.l[10] = c.z; r.l[11] = c.w;
        return r;
    }
    __device__ __forceinline__ static void store(uint32_t* p, const Fq& v) {
        uint4* q = reinterpret_cast<uint4*>(p);
        q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
        q[2] = make_uint4(v.l[8], v.l[9], v.l[10], v.l[11]);
    }
};
template <>
struct FieldIO<FqCall> {
    static constexpr int W = 12;
    __device__ __forceinline__ static FqCall load(const uint32_t* p) { return FqCall(FieldIO<Fq>::load(p)); }
    __device__ __forceinline__ static FqCall load_rw(const uint32_t* p) { return FqCall(FieldIO<Fq>::load_rw(p)); }
    __device__ __forceinline__ static void store(uint32_t* p, const FqCall& v) { FieldIO<Fq>::store(p, v); }
};
template <>
struct FieldIO<Fq2> {
    static constexpr int W = 24;
    __device__ __forceinline__ static Fq2 load(const uint32_t* p) { return Fq2{FieldIO<Fq>::load(p), FieldIO<Fq>::load(p + 12)}; }
    __device__ __forceinline__ static Fq2 load_rw(const uint32_t* p) {
        return Fq2{FieldIO<Fq>::load_rw(p), FieldIO<Fq>::load_rw(p + 12)};
    }
    __device__ __forceinline__ static void store(uint32_t* p, const Fq2& v) {
        FieldIO<Fq>::store(p, v.c0);
        FieldIO<Fq>::store(p + 12, v.c1);
    }
};
template <class F>
__device__ __forceinline__ XYZZ<F> load_point(const uint32_t* p) {
    constexpr int W = FieldIO<F>::W;
    XYZZ<F> r;
    r.x = FieldIO<F>::load_rw(p);
    r.y = FieldIO<F>::load_rw(p + W);
    r.zz = FieldIO<F>::load_rw(p + 2 * W);
    r.zzz = FieldIO<F>::load_rw(p + 3 * W);
    return r;
}
template <class F>
__device__ __forceinline__ void store_point(uint32_t* p, const XYZZ<F>& v) {
    constexpr int W = FieldIO<F>::W;
    FieldIO<F>::store(p, v.x);
    FieldIO<F>::store(p + W, v.y);
    FieldIO<F>::store(p + 2 * W, v.zz);
    FieldIO<F>::store(p + 3 * W, v.zzz);
}

}  // namespace czk
