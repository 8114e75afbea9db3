// Host/device portability macros shared by every header in csrc/.
#pragma once
#include <cstdint>
#include <cstddef>

#ifdef __CUDACC__
#define CZK_HD __host__ __device__ __forceinline__
#define CZK_D __device__ __forceinline__
#define CZK_HD_NOINLINE __host__ __device__ __noinline__
#else
#define CZK_HD inline
#define CZK_D inline
#define CZK_HD_NOINLINE inline
#endif
