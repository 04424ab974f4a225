// First conv layer of an image-like input (RGB / front view: C <= 3, K = 9*C <= 32) as a tcgen05 GEMM whose A operand
// is built in shared memory by the CTA itself (Network.conv on image_data, lib/networks/network.py:108-132 with
// MV3D_test.py:51-52).  The direct fp32 form (conv3x3_small_cin_kernel, layout_ops.cu) spends 1728 FMAs + 512 conversion
// instructions per pixel; here a thread gathers its pixel's 27 inputs once, splits them into bf16 hi/lo and writes one
// 64-byte K row per plane (128-byte swizzled pitch), six M128 x N64 x K16 MMAs (hi*hi + lo*hi + hi*lo over two k-steps,
// error ~2^-17 per product) replace the FMAs, and the epilogue renders the 64 channels of the pixel from TMEM straight
// into the consumer's operand format.  128 threads per CTA, 64 TMEM columns, 48 KB of shared memory: four CTAs per SM
// hide each other's gather / MMA / store latencies (no pipeline inside the CTA).
#include "common.cuh"
#include "ptx.cuh"

namespace mv3d {

using namespace ptx;

constexpr int kFlThreads = 128;
constexpr int kFlRowBytes = 128;                        // swizzled row pitch; K = 32 bf16 = the first 64 bytes of a row
constexpr int kFlAPlane = kFlThreads * kFlRowBytes;     // 16 KB
constexpr int kFlWPlane = 64 * kFlRowBytes;             // 8 KB
constexpr int kFlSmem = 2 * kFlAPlane + 2 * kFlWPlane + 1024;   // + alignment slack

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// 16-byte chunk c (0..7) of row r inside a SWIZZLE_128B K-major tile
__device__ __forceinline__ uint32_t swz128(int r, int c) { return (uint32_t)r * kFlRowBytes + (uint32_t)((c ^ (r & 7)) << 4); }

template <int FMT, int C>
__global__ void __launch_bounds__(kFlThreads, 4)
conv3x3_small_cin_mma_kernel(const float* __restrict__ in, int B, int H, int W, const float* __restrict__ w_hwio,
                             const float* __restrict__ bias, int relu, void* __restrict__ out_hi,
                             void* __restrict__ out_lo) {
    constexpr int K = 9 * C, N = 64;
    extern __shared__ unsigned char fl_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)fl_smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* a_hi = base;
    unsigned char* a_lo = base + kFlAPlane;
    unsigned char* w_hi = base + 2 * kFlAPlane;
    unsigned char* w_lo = w_hi + kFlWPlane;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float bias_s[N];
    const int tid = threadIdx.x, warp = tid >> 5;

    // weights: row n = output channel, K index = tap * C + c (HWIO flattened), zero beyond K; hi/lo split here
    for (int i = tid; i < N * 4; i += kFlThreads) {
        const int n = i >> 2, ch = i & 3;                    // 16-byte chunk = 8 K values
        __nv_bfloat16 h[8], l[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = ch * 8 + e;
            split_bf16(k < K ? w_hwio[k * N + n] : 0.f, h[e], l[e]);
        }
        *reinterpret_cast<uint4*>(w_hi + swz128(n, ch)) = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]),
                                                                     pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
        *reinterpret_cast<uint4*>(w_lo + swz128(n, ch)) = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]),
                                                                     pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
    }
    if (tid < N) bias_s[tid] = bias ? bias[tid] : 0.f;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(&tmem_slot, 64); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t idesc = make_idesc_bf16(128, N);
    const uint64_t da_hi = make_kmajor_desc(smem_u32(a_hi), 128), da_lo = make_kmajor_desc(smem_u32(a_lo), 128);
    const uint64_t dw_hi = make_kmajor_desc(smem_u32(w_hi), 128), dw_lo = make_kmajor_desc(smem_u32(w_lo), 128);

    const int Hp = H + 1, Wp = W + 1;
    const long long rows = (long long)B * Hp * Wp;
    const long long tiles = (rows + kFlThreads - 1) / kFlThreads;
    uint32_t phase = 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const long long pix = t * kFlThreads + tid;
        const int wp = (int)(pix % Wp);
        const long long r2 = pix / Wp;
        const int hp = (int)(r2 % Hp);
        const int b = (int)(r2 / Hp);
        const bool inside = pix < rows && wp > 0 && hp < H;
        // gather: x[k], k = (kh*3 + kw)*C + c at image (hp + kh - 1, wp - 1 + kw - 1), zero outside
        float x[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) x[k] = 0.f;
        if (inside) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int hh = hp + kh - 1;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int ww = wp + kw - 2;
                    if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
                        const float* px = in + (((long long)b * H + hh) * W + ww) * C;
#pragma unroll
                        for (int c = 0; c < C; ++c) x[(kh * 3 + kw) * C + c] = __ldg(px + c);
                    }
                }
            }
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            __nv_bfloat16 h[8], l[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split_bf16(x[ch * 8 + e], h[e], l[e]);
            *reinterpret_cast<uint4*>(a_hi + swz128(tid, ch)) = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]),
                                                                           pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
            *reinterpret_cast<uint4*>(a_lo + swz128(tid, ch)) = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]),
                                                                           pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
        }
        fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's async proxy
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 2; ++k) {  // two K = 16 steps (32 bytes each) of the 64 written bytes per row
                const uint64_t off = (uint64_t)(k * (32 >> 4));
                mma_bf16_ss(tmem, da_hi + off, dw_hi + off, idesc, k > 0 ? 1u : 0u);
                mma_bf16_ss(tmem, da_lo + off, dw_hi + off, idesc, 1u);
                mma_bf16_ss(tmem, da_hi + off, dw_lo + off, idesc, 1u);
            }
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1u;
        tc_fence_after();
        uint32_t v0[32], v1[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        tmem_ld_32x32(taddr, v0);
        tmem_ld_32x32(taddr + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncthreads();                   // accumulator read, A planes consumed: the next tile may overwrite both
        if (pix < rows) {
            float f[64];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                f[j] = __uint_as_float(v0[j]) + bias_s[j];
                f[32 + j] = __uint_as_float(v1[j]) + bias_s[32 + j];
            }
#pragma unroll
            for (int j = 0; j < 64; ++j) f[j] = inside ? (relu ? fmaxf(f[j], 0.f) : f[j]) : 0.f;
            if (FMT == MV3D_FMT_F16E5) {
                uint32_t h2[32], h8w[16], l8w[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    unsigned short ha, la, hb, lb;
                    split_f16e5_x2(f[4 * j], f[4 * j + 1], h2[2 * j], ha, la);
                    split_f16e5_x2(f[4 * j + 2], f[4 * j + 3], h2[2 * j + 1], hb, lb);
                    h8w[j] = (uint32_t)ha | ((uint32_t)hb << 16);
                    l8w[j] = (uint32_t)la | ((uint32_t)lb << 16);
                }
                unsigned char* ph = reinterpret_cast<unsigned char*>(out_hi) + pix * 128;   // 64 x fp16
                unsigned char* pl = reinterpret_cast<unsigned char*>(out_lo) + pix * 128;   // 64 x e5m2(h) | 64 x e5m2(residual)
#pragma unroll
                for (int q = 0; q < 4; ++q) st_global_v8(ph + 32 * q, h2 + 8 * q);
                st_global_v8(pl, h8w);
                st_global_v8(pl + 32, h8w + 8);
                st_global_v8(pl + 64, l8w);
                st_global_v8(pl + 96, l8w + 8);
            } else {
                uint32_t hw[32], lw[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    __nv_bfloat16 h0, l0, h1, l1;
                    split_bf16(f[2 * j], h0, l0);
                    split_bf16(f[2 * j + 1], h1, l1);
                    hw[j] = pack_bf16x2(h0, h1);
                    lw[j] = pack_bf16x2(l0, l1);
                }
                unsigned char* ph = reinterpret_cast<unsigned char*>(out_hi) + pix * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q) st_global_v8(ph + 32 * q, hw + 8 * q);
                if (out_lo) {
                    unsigned char* pl = reinterpret_cast<unsigned char*>(out_lo) + pix * 128;
#pragma unroll
                    for (int q = 0; q < 4; ++q) st_global_v8(pl + 32 * q, lw + 8 * q);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

// Called by mv3d_conv3x3_small_cin (layout_ops.cu) for Cout == c_pad == 64 and C <= 3.
int launch_small_cin_mma(const float* d_in, int B, int H, int W, int C, const float* d_w, const float* d_bias, int relu,
                         void* d_out_hi, void* d_out_lo, int fmt, cudaStream_t st) {
    const long long rows = (long long)B * (H + 1) * (W + 1);
    const long long tiles = (rows + kFlThreads - 1) / kFlThreads;
    int sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int grid = (int)(tiles < (long long)sms * 4 ? tiles : (long long)sms * 4);
#define MV3D_FL(FMT, CC)                                                                                              \
    do {                                                                                                              \
        static bool attr = false;                                                                                     \
        if (!attr) {                                                                                                  \
            cudaError_t e = cudaFuncSetAttribute(conv3x3_small_cin_mma_kernel<FMT, CC>,                               \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, kFlSmem);               \
            if (e != cudaSuccess) { set_last_cuda_error(e); return MV3D_ERR_LAUNCH; }                                 \
            attr = true;                                                                                              \
        }                                                                                                             \
        conv3x3_small_cin_mma_kernel<FMT, CC><<<grid, kFlThreads, kFlSmem, st>>>(d_in, B, H, W, d_w, d_bias, relu,    \
                                                                                 d_out_hi, d_out_lo);                 \
    } while (0)
    if (fmt == MV3D_FMT_F16E5) {
        if (C == 1) MV3D_FL(MV3D_FMT_F16E5, 1); else if (C == 2) MV3D_FL(MV3D_FMT_F16E5, 2); else MV3D_FL(MV3D_FMT_F16E5, 3);
    } else {
        if (C == 1) MV3D_FL(MV3D_FMT_BF16X2, 1); else if (C == 2) MV3D_FL(MV3D_FMT_BF16X2, 2); else MV3D_FL(MV3D_FMT_BF16X2, 3);
    }
#undef MV3D_FL
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

}  // namespace mv3d
