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
// out[i] = sum_p gathered[p*n + i]
cudaError_t fr_sum_parties(uint32_t* out, const uint32_t* gathered, size_t n, int parties, cudaStream_t st);
// sigma[i] = mac_share * x[i] - mac[i]     (spdz.rs:173-178; mac_share is 1 at the king and 0 elsewhere)
cudaError_t fr_spdz_sigma(uint32_t* sigma, const uint32_t* x, const uint32_t* mac, const uint64_t mac_share[4], size_t n,
                          cudaStream_t st);
// *flag |= 1 if any sum_p gathered[p*n+i] != 0
cudaError_t fr_check_zero_sum(const uint32_t* gathered, size_t n, int parties, uint32_t* flag, cudaStream_t st);
// d[i] = x[i] + tx   (open input of Beaver: share plus this party's triple share; tx is a per-party constant
// with the stub triple source)
cudaError_t fr_add_const(uint32_t* d, const uint32_t* x, const uint64_t tx[4], size_t n, cudaStream_t st);
// Beaver finish (share/field.rs:116-126) for one share component:
//   out[i] = tz - sx[i]*ty - oy[i]*tx + shift * sx[i]*oy[i]
// tx,ty,tz: this party's (constant) triple shares for this component; shift: 1 where the public product is
// added (king for the value share; mac_share for the MAC share), else 0.
cudaError_t fr_beaver_finish(uint32_t* out, const uint32_t* sx, const uint32_t* oy, const uint64_t tx[4],
                             const uint64_t ty[4], const uint64_t tz[4], const uint64_t shift[4], size_t n, cudaStream_t st);

}  // namespace czk
