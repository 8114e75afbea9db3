// Signed window decomposition of a canonical 253-bit scalar for the bucket method.
//
// The reference slices unsigned c-bit digits (algebra/ec/src/msm/variable_base.rs:50-57) and so
// needs 2^c - 1 buckets per window.  Bucket sums are not observable - only the final group element
// is - so the device uses the balanced form  sum_w d_w 2^(c w),  d_w in [-2^(c-1), 2^(c-1)],
// which halves the buckets (negating an affine point is free: y -> p - y).
#pragma once
#include "hd.cuh"

namespace czk {

// number of windows such that the top digit can never carry out (scalar < 2^253)
CZK_HD unsigned msm_num_windows(unsigned c) { return (254 + c - 1) / c; }

// Streaming form: call next(s, c, w) for w = 0, 1, 2, ... in order.
struct DigitCursor {
    uint32_t carry = 0;
    // s: 8 little-endian 32-bit words (any address space)
    CZK_HD int32_t next(const uint32_t* s, unsigned c, unsigned w) {
        const uint32_t half = 1u << (c - 1);
        const uint32_t mask = (1u << c) - 1u;  // c <= 24
        unsigned bit = w * c;
        unsigned i = bit >> 5, sh = bit & 31;
        uint32_t lo = i < 8 ? s[i] : 0u;
        uint32_t hi = i + 1 < 8 ? s[i + 1] : 0u;
        uint64_t v64 = (((uint64_t)hi << 32) | lo) >> sh;
        uint32_t v = ((uint32_t)v64 & mask) + carry;
        if (v > half) {
            carry = 1;
            return (int32_t)v - (int32_t)(1u << c);
        }
        carry = 0;
        return (int32_t)v;
    }
};

CZK_HD void signed_digits(const uint32_t* s, unsigned c, unsigned nwin, int32_t* out) {
    DigitCursor cur;
    for (unsigned w = 0; w < nwin; w++) out[w] = cur.next(s, c, w);
}

}  // namespace czk
