// Shared helpers for the mv3d_b200 kernels (status codes, launch checks, bf16 hi/lo split).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
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

// ----------------------------------------------------------------------------------------------------------------
// "f16e5" operand format (MV3D_FMT_F16E5): x ~= h + l / 4096 with h = fp16(x); next to the fp16 plane a byte plane
// holds, per 64-channel chunk, 64 x e5m2(h) followed by 64 x e5m2((x - h) * 4096)  (weights: residual first, then
// e5m2(w); their fp16 plane is fp16(4096 w)).  One 128-byte row of the byte plane is then the K-concatenation
// [A_h8 | A_l8] . [W_l8 ; W_h8] = 4096 (A_h W_l + A_l W_h): the two first-order correction terms of the split
// product as ONE fp8 MMA pass at twice the fp16 rate, accumulated on top of 4096 A_h W_h.  The GEMM epilogue
// multiplies by 2^-12.  Relative error per product ~2^-14.5 (bf16 hi/lo 3-pass: 2^-17; plain fp16: 2^-12).
// ----------------------------------------------------------------------------------------------------------------
constexpr float kF16E5Scale = 4096.f;

__device__ __forceinline__ uint8_t to_e5m2(float x) {
    return (uint8_t)__nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E5M2);
}
__device__ __forceinline__ float from_e5m2(uint8_t v) {  // e5m2 is the top byte of an fp16
    return __half2float(__ushort_as_half((unsigned short)(v << 8)));
}
// byte offset of channel c's e5m2(h) inside a byte-plane row; the residual sits 64 bytes further (activations)
__device__ __forceinline__ int f16e5_off(int c) { return ((c >> 6) << 7) + (c & 63); }

__device__ __forceinline__ void split_f16e5(float x, unsigned short& h, uint8_t& h8, uint8_t& l8) {
    x = fminf(fmaxf(x, -65504.f), 65504.f);
    const __half hh = __float2half_rn(x);
    const float hf = __half2float(hh);
    h = __half_as_ushort(hh);
    h8 = to_e5m2(hf);
    l8 = to_e5m2((x - hf) * kF16E5Scale);
}
__device__ __forceinline__ float join_f16e5(unsigned short h, uint8_t l8) {
    return __half2float(__ushort_as_half(h)) + from_e5m2(l8) * (1.f / kF16E5Scale);
}
// Two values at once with the packed conversion instructions (same roundings as split_f16e5): h2 = fp16x2 (x0 low),
// h8 / l8 = e5m2x2 of the fp16 values / of the scaled residuals.
__device__ __forceinline__ void split_f16e5_x2(float x0, float x1, uint32_t& h2, unsigned short& h8, unsigned short& l8) {
    x0 = fminf(fmaxf(x0, -65504.f), 65504.f);
    x1 = fminf(fmaxf(x1, -65504.f), 65504.f);
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(x1), "f"(x0));
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h2));
    asm("cvt.rn.satfinite.e5m2x2.f32 %0, %1, %2;" : "=h"(h8) : "f"(hf.y), "f"(hf.x));
    asm("cvt.rn.satfinite.e5m2x2.f32 %0, %1, %2;" : "=h"(l8) : "f"((x1 - hf.y) * kF16E5Scale), "f"((x0 - hf.x) * kF16E5Scale));
}
// weights: fp16 plane = fp16(4096 w), byte plane = [e5m2(4096 w - fp16 plane) | e5m2(w)]
__device__ __forceinline__ void split_f16e5_weight(float w, unsigned short& h, uint8_t& l8, uint8_t& h8) {
    const float ws = fminf(fmaxf(w * kF16E5Scale, -65504.f), 65504.f);
    const __half hh = __float2half_rn(ws);
    h = __half_as_ushort(hh);
    l8 = to_e5m2(ws - __half2float(hh));
    h8 = to_e5m2(w);
}

}  // namespace mv3d
