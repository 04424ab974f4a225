// Shared helpers for the mv3d_b200 kernels (status codes, launch checks, bf16 hi/lo split).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mv3d_b200.h"

#define MV3D_CHECK_LAUNCH()                                            \
    do {                                                               \
        cudaError_t e__ = cudaGetLastError();                          \
        if (e__ != cudaSuccess) { mv3d::set_last_cuda_error(e__); return MV3D_ERR_LAUNCH; } \
    } while (0)

#define MV3D_REQUIRE(cond)                 \
    do {                                   \
        if (!(cond)) return MV3D_ERR_ARG;  \
    } while (0)

namespace mv3d {

void set_last_cuda_error(cudaError_t e);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits kept across the pair.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

}  // namespace mv3d
