// First conv layer of an image-like input (RGB / front view: C <= 3, K = 9*C <= 32) as a tcgen05 GEMM whose A operand
// is built in shared memory by the CTA itself (Network.conv on image_data, lib/networks/network.py:108-132 with
// MV3D_test.py:51-52).  The direct fp32 form (conv3x3_small_cin_kernel, layout_ops.cu) spends 1728 FMAs + 512 conversion
// instructions per pixel; here a thread gathers its pixel's 27 inputs once, splits them into bf16 hi/lo and writes one
// 64-byte K row per plane (SWIZZLE_64B tiles), six M128 x N64 x K16 MMAs (hi*hi + lo*hi + hi*lo over two k-steps,
// error ~2^-17 per product) replace the FMAs, and the epilogue renders the 64 channels of the pixel from TMEM straight
// into the consumer's operand format.  128 threads per CTA, 64 TMEM columns, 25 KB of shared memory: five CTAs per SM
// hide each other's gather / MMA / store latencies (no pipeline inside the CTA).
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace mv3d {

using namespace ptx;

#ifndef MV3D_FL_MINBLOCKS
#define MV3D_FL_MINBLOCKS 3
#endif
constexpr int kFlThreads = 128;
constexpr int kFlRowBytes = 64;                         // K = 32 bf16 per row, SWIZZLE_64B K-major tiles (8-row atoms of 512 B)
constexpr int kFlAPlane = kFlThreads * kFlRowBytes;     // 8 KB
constexpr int kFlWPlane = 64 * kFlRowBytes;             // 4 KB
constexpr int kFlStage = kFlThreads * 128;              // one output plane of a tile: 128 pixels x 128 bytes (SWIZZLE_128B)
constexpr int kFlSmem = 2 * kFlStage + 2 * kFlAPlane + 2 * kFlWPlane + 1024;   // + alignment slack

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// 16-byte chunk c (0..3) of row r inside a SWIZZLE_64B K-major tile: address bits [4:5] ^= bits [7:8]
__device__ __forceinline__ uint32_t swz64(int r, int c) { return (uint32_t)r * kFlRowBytes + (uint32_t)((c ^ ((r >> 1) & 3)) << 4); }

template <int FMT, int C>
__global__ void __launch_bounds__(kFlThreads, MV3D_FL_MINBLOCKS)
conv3x3_small_cin_mma_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                             const float* __restrict__ in, int B, int H, int W, const float* __restrict__ w_hwio,
                             const float* __restrict__ bias, int relu, int has_lo) {
    constexpr int K = 9 * C, N = 64;
    extern __shared__ unsigned char fl_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)fl_smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* st_hi = base;                       // output staging: the tile's two planes, 1024-byte aligned
    unsigned char* st_lo = base + kFlStage;
    unsigned char* a_hi = base + 2 * kFlStage;
    unsigned char* a_lo = a_hi + kFlAPlane;
    unsigned char* w_hi = a_lo + kFlAPlane;
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
        *reinterpret_cast<uint4*>(w_hi + swz64(n, ch)) = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]),
                                                                     pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
        *reinterpret_cast<uint4*>(w_lo + swz64(n, ch)) = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]),
                                                                     pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
    }
    if (tid < N) bias_s[tid] = bias ? bias[tid] : 0.f;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); prefetch_tensormap(&map_hi); prefetch_tensormap(&map_lo); }
    if (warp == 0) { tmem_alloc(&tmem_slot, 64); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t idesc = make_idesc_bf16(128, N);
    const uint64_t da_hi = make_kmajor_desc(smem_u32(a_hi), kFlRowBytes), da_lo = make_kmajor_desc(smem_u32(a_lo), kFlRowBytes);
    const uint64_t dw_hi = make_kmajor_desc(smem_u32(w_hi), kFlRowBytes), dw_lo = make_kmajor_desc(smem_u32(w_lo), kFlRowBytes);

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
        // hi/lo split two values per conversion instruction (= split_bf16 on each): the kernel's time is the F2FP pipe
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            uint32_t h2[4], l2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float x0 = x[ch * 8 + 2 * e], x1 = x[ch * 8 + 2 * e + 1];
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h2[e]) : "f"(x1), "f"(x0));
                const float r0 = x0 - __uint_as_float(h2[e] << 16), r1 = x1 - __uint_as_float(h2[e] & 0xFFFF0000u);
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l2[e]) : "f"(r1), "f"(r0));
            }
            *reinterpret_cast<uint4*>(a_hi + swz64(tid, ch)) = make_uint4(h2[0], h2[1], h2[2], h2[3]);
            *reinterpret_cast<uint4*>(a_lo + swz64(tid, ch)) = make_uint4(l2[0], l2[1], l2[2], l2[3]);
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
        if (tid == 0) bulk_wait_group_read0();   // the previous tile's stores have read the staging planes
        __syncthreads();                   // accumulator read, A planes consumed, staging free
        {
            float f[64];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                f[j] = __uint_as_float(v0[j]) + bias_s[j];
                f[32 + j] = __uint_as_float(v1[j]) + bias_s[32 + j];
            }
#pragma unroll
            for (int j = 0; j < 64; ++j) f[j] = inside ? (relu ? fmaxf(f[j], 0.f) : f[j]) : 0.f;
            // The tile's output planes are staged in shared memory (row = pixel, 128 B, SWIZZLE_128B) and leave as ONE
            // tensor-map store per plane: a per-thread row store touches 32 different lines per warp instruction.
            uint32_t hw[32], lw[32];
            if (FMT == MV3D_FMT_F16E5) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {   // fp16 plane | byte plane = 64 x e5m2(h), 64 x e5m2(residual)
                    unsigned short ha, la, hb, lb;
                    split_f16e5_x2(f[4 * j], f[4 * j + 1], hw[2 * j], ha, la);
                    split_f16e5_x2(f[4 * j + 2], f[4 * j + 3], hw[2 * j + 1], hb, lb);
                    lw[j] = (uint32_t)ha | ((uint32_t)hb << 16);
                    lw[16 + j] = (uint32_t)la | ((uint32_t)lb << 16);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float x0 = f[2 * j], x1 = f[2 * j + 1];
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hw[j]) : "f"(x1), "f"(x0));
                    const float r0 = x0 - __uint_as_float(hw[j] << 16), r1 = x1 - __uint_as_float(hw[j] & 0xFFFF0000u);
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lw[j]) : "f"(r1), "f"(r0));
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint32_t o = (uint32_t)tid * 128u + (uint32_t)((c ^ (tid & 7)) << 4);
                *reinterpret_cast<uint4*>(st_hi + o) = make_uint4(hw[4 * c], hw[4 * c + 1], hw[4 * c + 2], hw[4 * c + 3]);
                *reinterpret_cast<uint4*>(st_lo + o) = make_uint4(lw[4 * c], lw[4 * c + 1], lw[4 * c + 2], lw[4 * c + 3]);
            }
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {   // rows past the end of the map are clipped by the hardware
            const long long row0 = t * kFlThreads;
            tma_store_2d(st_hi, &map_hi, 0, (int32_t)row0);
            if (has_lo) tma_store_2d(st_lo, &map_lo, 0, (int32_t)row0);
            bulk_commit_group();
        }
    }
    if (tid == 0) bulk_wait_group0();
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
    MV3D_REQUIRE(rows < (1LL << 31));
    // output planes as (pixels, 64 x 16 bit) matrices: both renderings have 128-byte rows per pixel at c_pad = 64
    CUtensorMap m_hi, m_lo;
    int rc;
    if ((rc = make_map_2d(&m_hi, d_out_hi, (uint64_t)rows, 64, kFlThreads, 64)) != MV3D_OK) return rc;
    if ((rc = make_map_2d(&m_lo, d_out_lo ? d_out_lo : d_out_hi, (uint64_t)rows, 64, kFlThreads, 64)) != MV3D_OK) return rc;
    const int grid = (int)(tiles < (long long)sms * MV3D_FL_MINBLOCKS ? tiles : (long long)sms * MV3D_FL_MINBLOCKS);
#define MV3D_FL(FMT, CC)                                                                                              \
    do {                                                                                                              \
        static bool attr = false;                                                                                     \
        if (!attr) {                                                                                                  \
            cudaError_t e = cudaFuncSetAttribute(conv3x3_small_cin_mma_kernel<FMT, CC>,                               \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, kFlSmem);               \
            if (e != cudaSuccess) { set_last_cuda_error(e); return MV3D_ERR_LAUNCH; }                                 \
            attr = true;                                                                                              \
        }                                                                                                             \
        conv3x3_small_cin_mma_kernel<FMT, CC><<<grid, kFlThreads, kFlSmem, st>>>(m_hi, m_lo, d_in, B, H, W, d_w,      \
                                                                                 d_bias, relu, d_out_lo ? 1 : 0);     \
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
