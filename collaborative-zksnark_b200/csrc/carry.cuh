// 32-bit carry-chain primitives.
//
// On the device each one is a single PTX instruction using the condition-code
// register (add.cc / addc / mad.lo.cc / madc.hi.cc ...); ptxas fuses a
// (mad[c].lo.cc, madc.hi.cc) pair on the same operands into one
// IMAD.WIDE.U32[.X] with carry in/out, which is the instruction the whole
// field layer is built to issue.  Every statement is `asm volatile` because the
// carry flag is invisible to the compiler: without it two identical-looking
// statements could be merged or reordered.
//
// On the host (g++ build used by tests/test_device_math_emulation.py and by
// nothing in the product path) the same functions emulate the flag in a
// thread-local, so the exact limb schedule that runs on the GPU can be checked
// on a CPU-only machine.
#pragma once
#include "hd.cuh"

namespace czk {

#ifndef __CUDA_ARCH__
struct CarryFlag {
    static uint32_t& cf() {
        static thread_local uint32_t v = 0;
        return v;
    }
};
#endif

CZK_HD uint32_t add_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    uint64_t t = (uint64_t)a + b;
    CarryFlag::cf() = (uint32_t)(t >> 32);
    return (uint32_t)t;
#endif
}
CZK_HD uint32_t addc_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    uint64_t t = (uint64_t)a + b + CarryFlag::cf();
    CarryFlag::cf() = (uint32_t)(t >> 32);
    return (uint32_t)t;
#endif
}
CZK_HD uint32_t addc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    return a + b + CarryFlag::cf();
#endif
}
CZK_HD uint32_t sub_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    // PTX models the borrow as CF=1 meaning "borrow occurred" for subc
    uint64_t t = (uint64_t)a - b;
    CarryFlag::cf() = (uint32_t)(t >> 63);
    return (uint32_t)t;
#endif
}
CZK_HD uint32_t subc_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    uint64_t t = (uint64_t)a - b - CarryFlag::cf();
    CarryFlag::cf() = (uint32_t)(t >> 63);
    return (uint32_t)t;
#endif
}
CZK_HD uint32_t subc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    return a - b - CarryFlag::cf();
#endif
}
CZK_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
CZK_HD uint32_t mul_hi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
// r = lo(a*b) + c, sets CF
CZK_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    uint64_t t = (uint64_t)(uint32_t)(a * b) + c;
    CarryFlag::cf() = (uint32_t)(t >> 32);
    return (uint32_t)t;
#endif
}
CZK_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    uint64_t t = (uint64_t)(uint32_t)(a * b) + c + CarryFlag::cf();
    CarryFlag::cf() = (uint32_t)(t >> 32);
    return (uint32_t)t;
#endif
}
CZK_HD uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    uint64_t t = (uint64_t)mul_hi(a, b) + c;
    CarryFlag::cf() = (uint32_t)(t >> 32);
    return (uint32_t)t;
#endif
}
CZK_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    uint64_t t = (uint64_t)mul_hi(a, b) + c + CarryFlag::cf();
    CarryFlag::cf() = (uint32_t)(t >> 32);
    return (uint32_t)t;
#endif
}
CZK_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    return mul_hi(a, b) + c + CarryFlag::cf();
#endif
}

}  // namespace czk
