// Streaming (HBM-bound) elementwise kernels on Fr vectors: the pointwise steps of the witness map
// (mpc-snarks/src/groth/r1cs_to_qap.rs:92,105-109), share opening (mpc-algebra/src/share/add.rs:121-125,
// spdz.rs:166-185) and the local half of Beaver multiplication (share/field.rs:97-127).
#pragma once
#include <cuda_runtime.h>
#include "fp.cuh"

namespace czk {

struct FrConst {
    uint64_t v[4];
};

enum FrBinOp { FR_ADD = 0, FR_SUB = 1, FR_MUL = 2 };
cudaError_t fr_binop(uint32_t* a, const uint32_t* b, size_t n, FrBinOp op, cudaStream_t st);     // a = a op b
cudaError_t fr_scale(uint32_t* a, const uint64_t c[4], size_t n, cudaStream_t st);                // a *= c
// a[i] *= lo[i & (2^lo_log - 1)] * hi[i >> lo_log]   (two-level power table, any n)
cudaError_t fr_scale_by_tables(uint32_t* a, const uint32_t* lo, const uint32_t* hi, int lo_log, size_t n, cudaStream_t st);
// d[i] = x[i] + tx   (a share plus a per-party constant)
cudaError_t fr_add_const(uint32_t* d, const uint32_t* x, const uint64_t tx[4], size_t n, cudaStream_t st);
}  // namespace czk
