// Layout / elementwise kernels around the tcgen05 GEMM: weight packing, PAD-layout conversion,
// 2x2 max-pool on the PAD layout, pair-softmax, split-K epilogue.  All HBM-bound, one pass each.
#include <stdlib.h>

#include "common.cuh"

namespace mv3d {

// HWIO fp32 (taps, cin, cout) -> bf16 hi/lo (cout, taps*cin_pad), zero channel padding.
__global__ void pack_weights_kernel(const float* __restrict__ w, int taps, int cin, int cout, int cin_pad,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const long long total = (long long)cout * taps * cin_pad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cin_pad);
        const long long r = i / cin_pad;
        const int t = (int)(r % taps);
        const int n = (int)(r / taps);
        float x = 0.f;
        if (c < cin) x = w[((long long)t * cin + c) * cout + n];
        __nv_bfloat16 h, l;
        split_bf16(x, h, l);
        hi[i] = h;
        if (lo) lo[i] = l;
    }
}

// Same packing as a 32x32 shared-memory transpose: reads run along cout (contiguous in HWIO), writes along the packed K
// axis (contiguous in the output) -- both coalesced.  Used when cout >= 32 (every layer but the tiny heads).
__global__ void pack_weights_tiled_kernel(const float* __restrict__ w, int taps, int cin, int cout, int cin_pad,
                                          __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    __shared__ float tile[32][33];
    const int kdim = taps * cin_pad;
    const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {       // j: packed-K row of the tile, threadIdx.x: cout lane
        const int k = k0 + j, n = n0 + threadIdx.x;
        float x = 0.f;
        if (k < kdim && n < cout) {
            const int t = k / cin_pad, c = k - t * cin_pad;
            if (c < cin) x = w[((long long)t * cin + c) * cout + n];
        }
        tile[j][threadIdx.x] = x;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {       // j: cout row of the output, threadIdx.x: packed-K lane
        const int n = n0 + j, k = k0 + threadIdx.x;
        if (n < cout && k < kdim) {
            __nv_bfloat16 h, l;
            split_bf16(tile[threadIdx.x][j], h, l);
            hi[(long long)n * kdim + k] = h;
            if (lo) lo[(long long)n * kdim + k] = l;
        }
    }
}

// 64(K) x 64(cout) tiles, 256 threads: reads run along cout (256 B per k row), every thread then emits EIGHT consecutive
// packed-K elements of one output row as one 16-byte store per plane (8 lanes cover 128 contiguous bytes of a row).
// The packers run once per weight version -- once per training step for all 143 M parameters.
__global__ void __launch_bounds__(256)
pack_weights_tiled64_kernel(const float* __restrict__ w, int taps, int cin, int cout, int cin_pad,
                            __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    __shared__ float tile[64][65];
    const int kdim = taps * cin_pad;
    const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
#pragma unroll 4
    for (int j = ty; j < 64; j += 4) {
        const int k = k0 + j, n = n0 + tx;
        float x = 0.f;
        if (k < kdim && n < cout) {
            const int t = k / cin_pad, c = k - t * cin_pad;
            if (c < cin) x = __ldg(w + ((long long)t * cin + c) * cout + n);
        }
        tile[j][tx] = x;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int item = threadIdx.x + it * 256;
        const int kg = item & 7, nn = item >> 3;
        const int n = n0 + nn, k = k0 + kg * 8;
        if (n >= cout || k >= kdim) continue;       // kdim % 8 == 0 (cin_pad is a multiple of 16)
        uint32_t ph[4], pl[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(tile[kg * 8 + 2 * e][nn], h0, l0);
            split_bf16(tile[kg * 8 + 2 * e + 1][nn], h1, l1);
            ph[e] = uint32_t(__bfloat16_as_ushort(h0)) | (uint32_t(__bfloat16_as_ushort(h1)) << 16);
            pl[e] = uint32_t(__bfloat16_as_ushort(l0)) | (uint32_t(__bfloat16_as_ushort(l1)) << 16);
        }
        const long long o = (long long)n * kdim + k;
        *reinterpret_cast<uint4*>(hi + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        if (lo) *reinterpret_cast<uint4*>(lo + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
}

// (B,H,W,C) fp32 -> PAD (B,H+1,W+1,c_pad) bf16 hi/lo.  One thread per 8 output channels (16-byte stores).
__global__ void pad_nhwc_kernel(const float* __restrict__ in, int B, int H, int W, int C, int c_pad,
                                __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const int Hp = H + 1, Wp = W + 1, cv = c_pad / 8;
    const long long total = (long long)B * Hp * Wp * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cv);
        long long r = i / cv;
        const int wp = (int)(r % Wp);
        r /= Wp;
        const int hp = (int)(r % Hp);
        const int b = (int)(r / Hp);
        __nv_bfloat16 vh[8], vl[8];
        const bool inside = wp > 0 && hp < H;
        const float* src = in + (((long long)b * H + hp) * W + (wp - 1)) * C;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = c8 * 8 + e;
            split_bf16((inside && c < C) ? src[c] : 0.f, vh[e], vl[e]);
        }
        *reinterpret_cast<uint4*>(hi + i * 8) = *reinterpret_cast<uint4*>(vh);
        if (lo) *reinterpret_cast<uint4*>(lo + i * 8) = *reinterpret_cast<uint4*>(vl);
    }
}

// First-layer im2col for tiny channel counts (RGB / front view, C = 3): dense fp32 (B,H,W,C) -> PAD rows
// (B,H+1,W+1,k_pad) whose K index is tap*C + c (tap = kh*3 + kw, SAME zero padding), so that conv1_1 becomes ONE K=32
// GEMM instead of nine K=16 taps that are 13/16 zero padding.  Halo rows and k >= 9*C are zero.
__global__ void im2col3x3_pad_kernel(const float* __restrict__ in, int B, int H, int W, int C, int k_pad,
                                     __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const int Hp = H + 1, Wp = W + 1, kv = k_pad / 8;
    const long long total = (long long)B * Hp * Wp * kv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int k8 = (int)(i % kv);
        long long r = i / kv;
        const int wp = (int)(r % Wp);
        r /= Wp;
        const int hp = (int)(r % Hp);
        const int b = (int)(r / Hp);
        __nv_bfloat16 vh[8], vl[8];
        const bool inside = wp > 0 && hp < H;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = k8 * 8 + e;
            float x = 0.f;
            if (inside && k < 9 * C) {
                const int tap = k / C, c = k - tap * C;
                const int hh = hp + tap / 3 - 1, ww = (wp - 1) + tap % 3 - 1;
                if (hh >= 0 && hh < H && ww >= 0 && ww < W) x = in[(((long long)b * H + hh) * W + ww) * C + c];
            }
            split_bf16(x, vh[e], vl[e]);
        }
        *reinterpret_cast<uint4*>(hi + i * 8) = *reinterpret_cast<uint4*>(vh);
        if (lo) *reinterpret_cast<uint4*>(lo + i * 8) = *reinterpret_cast<uint4*>(vl);
    }
}

// First conv layer of an image-like input with a handful of channels (RGB / front view: Cin = 3, K = 27): 1.6 of the
// frame's 920 GFLOP, but as a tensor-core GEMM it is an im2col pass plus a K = 32 GEMM whose time is all epilogue
// (82 us together at 375x1242).  Here it is ONE direct kernel, register-tiled: a thread owns 4 consecutive pixels of a
// row x 8 output channels (32 fp32 accumulators); the 3 x 6 x C input window is loaded once into registers, every
// weight vector (shared memory, [tap*C + c][cout]) is used for 4 pixels (1 LDS.128 per 16 FMAs); fp32 FMA accumulation
// in tap order, + bias, ReLU, then the consumer's operand rendering (f16e5 or bf16 hi/lo) with 16-byte stores straight
// into the PAD layout, halo pixels zero.  Bound by its 4 B/element output.
constexpr int kSmallPx = 4;

template <int FMT, int C>
__global__ void __launch_bounds__(256)
conv3x3_small_cin_kernel(const float* __restrict__ in, int B, int H, int W, const float* __restrict__ w_hwio,
                         const float* __restrict__ bias, int Cout, int relu, void* __restrict__ out_hi,
                         void* __restrict__ out_lo, int c_pad) {
    extern __shared__ float wsm[];                 // [9*C][Cout] + bias[Cout]
    constexpr int K = 9 * C;
    for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) wsm[i] = w_hwio[i];
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) wsm[K * Cout + i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const float* bsm = wsm + K * Cout;
    const int Hp = H + 1, Wp = W + 1, groups = c_pad / 8;
    const int wq = (Wp + kSmallPx - 1) / kSmallPx;              // 4-pixel segments per PAD row (halo column included)
    const long long total = (long long)B * Hp * wq * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int gq = (int)(i % groups);
        long long r = i / groups;
        const int seg = (int)(r % wq);
        r /= wq;
        const int hp = (int)(r % Hp);
        const int b = (int)(r / Hp);
        const int co = gq * 8;
        const int wp0 = seg * kSmallPx;                         // first PAD column of the segment; pixel w = wp - 1
        float acc[kSmallPx][8];
#pragma unroll
        for (int p = 0; p < kSmallPx; ++p)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[p][e] = 0.f;
        const bool row_inside = hp < H && co < Cout;
        if (row_inside) {
            // input window: rows hp-1..hp+1, image columns (wp0-1)-1 .. (wp0-1)+4, zero outside the image
            float x[3][kSmallPx + 2][C];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int hh = hp + kh - 1;
#pragma unroll
                for (int q = 0; q < kSmallPx + 2; ++q) {
                    const int ww = wp0 - 2 + q;
                    const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
                    const float* px = in + (((long long)b * H + (ok ? hh : 0)) * W + (ok ? ww : 0)) * C;
#pragma unroll
                    for (int c = 0; c < C; ++c) x[kh][q][c] = ok ? __ldg(px + c) : 0.f;
                }
            }
#pragma unroll
            for (int p = 0; p < kSmallPx; ++p)
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[p][e] = bsm[co + e];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float* wt = wsm + (size_t)((kh * 3 + kw) * C + c) * Cout + co;
                        const float4 w0 = *reinterpret_cast<const float4*>(wt);
                        const float4 w1 = *reinterpret_cast<const float4*>(wt + 4);
                        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                        for (int p = 0; p < kSmallPx; ++p) {
                            const float xv = x[kh][p + kw][c];
#pragma unroll
                            for (int e = 0; e < 8; ++e) acc[p][e] = __fmaf_rn(xv, wv[e], acc[p][e]);
                        }
                    }
        }
#pragma unroll
        for (int p = 0; p < kSmallPx; ++p) {
            const int wp = wp0 + p;
            if (wp >= Wp) break;
            const bool inside = row_inside && wp > 0;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = inside ? (relu ? fmaxf(acc[p][e], 0.f) : acc[p][e]) : 0.f;
            const long long pix = ((long long)b * Hp + hp) * Wp + wp;
            if (FMT == MV3D_FMT_F16E5) {
                uint32_t h2[4];
                unsigned short h8[4], l8[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split_f16e5_x2(v[2 * e], v[2 * e + 1], h2[e], h8[e], l8[e]);
                *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(out_hi) + pix * c_pad + co) = make_uint4(h2[0], h2[1], h2[2], h2[3]);
                uint8_t* row = reinterpret_cast<uint8_t*>(out_lo) + pix * c_pad * 2 + f16e5_off(co);
                *reinterpret_cast<uint2*>(row) = make_uint2(h8[0] | ((uint32_t)h8[1] << 16), h8[2] | ((uint32_t)h8[3] << 16));
                *reinterpret_cast<uint2*>(row + 64) = make_uint2(l8[0] | ((uint32_t)l8[1] << 16), l8[2] | ((uint32_t)l8[3] << 16));
            } else {
                __nv_bfloat16 vh[8], vl[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split_bf16(v[e], vh[e], vl[e]);
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out_hi) + pix * c_pad + co) = *reinterpret_cast<uint4*>(vh);
                if (out_lo) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out_lo) + pix * c_pad + co) = *reinterpret_cast<uint4*>(vl);
            }
        }
    }
}

__global__ void unpad_nhwc_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int B,
                                  int H, int W, int C, int c_pad, float* __restrict__ out) {
    const int Hp = H + 1, Wp = W + 1;
    const long long total = (long long)B * H * W * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long r = i / C;
        const int w = (int)(r % W);
        r /= W;
        const int h = (int)(r % H);
        const int b = (int)(r / H);
        const long long j = (((long long)b * Hp + h) * Wp + (w + 1)) * c_pad + c;
        float x = __bfloat162float(hi[j]);
        if (lo) x += __bfloat162float(lo[j]);
        out[i] = x;
    }
}

// 2x2/2 VALID max-pool, PAD in (B,H+1,W+1,c) -> PAD out (B,H/2+1,W/2+1,c).  One thread handles 8 channels
// (16-byte vectors).  The max is taken on hi+lo and the winner's (hi,lo) pair is copied, so it is exact.
__global__ void maxpool2x2_pad_kernel(const __nv_bfloat16* __restrict__ ih, const __nv_bfloat16* __restrict__ il,
                                      int B, int H, int W, int c_pad, __nv_bfloat16* __restrict__ oh,
                                      __nv_bfloat16* __restrict__ ol) {
    const int Ho = H / 2, Wo = W / 2;
    const int Hp = H + 1, Wp = W + 1, Hop = Ho + 1, Wop = Wo + 1;
    const int cv = c_pad / 8;
    const long long total = (long long)B * Hop * Wop * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cv);
        long long r = i / cv;
        const int wop = (int)(r % Wop);
        r /= Wop;
        const int hop = (int)(r % Hop);
        const int b = (int)(r / Hop);
        const long long o = ((((long long)b * Hop + hop) * Wop + wop) * c_pad) + c8 * 8;
        uint4 rh = make_uint4(0, 0, 0, 0), rl = make_uint4(0, 0, 0, 0);
        if (wop > 0 && hop < Ho) {
            const int h0 = hop * 2, w0 = (wop - 1) * 2;
            float best[8];
            __nv_bfloat16 bh[8], bl[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { best[e] = -3.4e38f; bh[e] = __float2bfloat16_rn(0.f); bl[e] = bh[e]; }
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const long long j = ((((long long)b * Hp + h0 + dy) * Wp + (w0 + dx + 1)) * c_pad) + c8 * 8;
                    uint4 vh = *reinterpret_cast<const uint4*>(ih + j);
                    uint4 vl = il ? *reinterpret_cast<const uint4*>(il + j) : make_uint4(0, 0, 0, 0);
                    const __nv_bfloat16* ph = reinterpret_cast<const __nv_bfloat16*>(&vh);
                    const __nv_bfloat16* pl = reinterpret_cast<const __nv_bfloat16*>(&vl);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float x = __bfloat162float(ph[e]) + __bfloat162float(pl[e]);
                        if (x > best[e]) { best[e] = x; bh[e] = ph[e]; bl[e] = pl[e]; }
                    }
                }
            rh = *reinterpret_cast<uint4*>(bh);
            rl = *reinterpret_cast<uint4*>(bl);
        }
        *reinterpret_cast<uint4*>(oh + o) = rh;
        if (ol) *reinterpret_cast<uint4*>(ol + o) = rl;
    }
}

// softmax over adjacent channel pairs: out[r, 2a+k] = exp(x_k - m) / (exp(x_0 - m) + exp(x_1 - m)).
__global__ void softmax_pairs_kernel(const float* __restrict__ in, int rows, int ld_in, int n_pairs,
                                     float* __restrict__ out, int ld_out) {
    const long long total = (long long)rows * n_pairs;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(i % n_pairs);
        const long long r = i / n_pairs;
        const float x0 = in[r * ld_in + 2 * a], x1 = in[r * ld_in + 2 * a + 1];
        const float m = fmaxf(x0, x1);
        const float e0 = expf(x0 - m), e1 = expf(x1 - m);
        const float s = e0 + e1;
        out[r * ld_out + 2 * a] = e0 / s;
        out[r * ld_out + 2 * a + 1] = e1 / s;
    }
}

__global__ void bias_act_kernel(const float* __restrict__ acc, int M, int N, int ld_acc,
                                const float* __restrict__ bias, int relu, __nv_bfloat16* __restrict__ hi,
                                __nv_bfloat16* __restrict__ lo, int ld_out, float* __restrict__ of32, int ld_f32) {
    const long long total = (long long)M * N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i % N);
        const long long m = i / N;
        float x = acc[m * ld_acc + n];
        if (bias) x += bias[n];
        if (relu) x = fmaxf(x, 0.f);
        if (hi) {
            __nv_bfloat16 h, l;
            split_bf16(x, h, l);
            hi[m * ld_out + n] = h;
            if (lo) lo[m * ld_out + n] = l;
        }
        if (of32) of32[m * ld_f32 + n] = x;
    }
}

// ----------------------------------------------------------------------------------------------------------------
// MV3D_FMT_F16E5 renderings of the same layout kernels (format: common.cuh).  `hi` is the fp16 plane, `lo` the byte
// plane; both have the pitch of the bf16 planes (c_pad 16-bit units per row), c_pad % 64 == 0.
// ----------------------------------------------------------------------------------------------------------------
__global__ void pack_weights_f16e5_kernel(const float* __restrict__ w, int taps, int cin, int cout, int cin_pad,
                                          unsigned short* __restrict__ hi, uint8_t* __restrict__ lo) {
    __shared__ float tile[32][33];
    const int kdim = taps * cin_pad;
    const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int k = k0 + j, n = n0 + threadIdx.x;
        float x = 0.f;
        if (k < kdim && n < cout) {
            const int t = k / cin_pad, c = k - t * cin_pad;
            if (c < cin) x = w[((long long)t * cin + c) * cout + n];
        }
        tile[j][threadIdx.x] = x;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int n = n0 + j, k = k0 + threadIdx.x;
        if (n < cout && k < kdim) {
            unsigned short h;
            uint8_t l8, h8;
            split_f16e5_weight(tile[threadIdx.x][j], h, l8, h8);
            hi[(long long)n * kdim + k] = h;
            uint8_t* row = lo + (long long)n * kdim * 2;
            row[f16e5_off(k)] = l8;        // residual first ...
            row[f16e5_off(k) + 64] = h8;   // ... then e5m2(w): pairs with the activation row [e5m2(h) | residual]
        }
    }
}

__global__ void pad_nhwc_f16e5_kernel(const float* __restrict__ in, int B, int H, int W, int C, int c_pad,
                                      unsigned short* __restrict__ hi, uint8_t* __restrict__ lo) {
    const int Hp = H + 1, Wp = W + 1, cv = c_pad / 8;
    const long long total = (long long)B * Hp * Wp * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cv);
        long long r = i / cv;
        const long long row = r;
        const int wp = (int)(r % Wp);
        r /= Wp;
        const int hp = (int)(r % Hp);
        const int b = (int)(r / Hp);
        const bool inside = wp > 0 && hp < H;
        const float* src = in + (((long long)b * H + hp) * W + (wp - 1)) * C;
        unsigned short vh[8];
        uint8_t a8[8], b8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = c8 * 8 + e;
            split_f16e5((inside && c < C) ? src[c] : 0.f, vh[e], a8[e], b8[e]);
        }
        *reinterpret_cast<uint4*>(hi + i * 8) = *reinterpret_cast<uint4*>(vh);
        uint8_t* ob = lo + row * c_pad * 2 + f16e5_off(c8 * 8);
        *reinterpret_cast<uint2*>(ob) = *reinterpret_cast<uint2*>(a8);
        *reinterpret_cast<uint2*>(ob + 64) = *reinterpret_cast<uint2*>(b8);
    }
}

__global__ void unpad_nhwc_f16e5_kernel(const unsigned short* __restrict__ hi, const uint8_t* __restrict__ lo, int B,
                                        int H, int W, int C, int c_pad, float* __restrict__ out) {
    const int Hp = H + 1, Wp = W + 1;
    const long long total = (long long)B * H * W * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long r = i / C;
        const int w = (int)(r % W);
        r /= W;
        const int h = (int)(r % H);
        const int b = (int)(r / H);
        const long long row = ((long long)b * Hp + h) * Wp + (w + 1);
        out[i] = join_f16e5(hi[row * c_pad + c], lo[row * c_pad * 2 + f16e5_off(c) + 64]);
    }
}

// 2x2/2 VALID max-pool: the winner by decoded value, its (fp16, e5m2, e5m2) triple copied unchanged.
__global__ void maxpool2x2_pad_f16e5_kernel(const unsigned short* __restrict__ ih, const uint8_t* __restrict__ il,
                                            int B, int H, int W, int c_pad, unsigned short* __restrict__ oh,
                                            uint8_t* __restrict__ ol) {
    const int Ho = H / 2, Wo = W / 2;
    const int Hp = H + 1, Wp = W + 1, Hop = Ho + 1, Wop = Wo + 1;
    const int cv = c_pad / 8;
    const long long total = (long long)B * Hop * Wop * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cv);
        long long r = i / cv;
        const long long orow = r;
        const int wop = (int)(r % Wop);
        r /= Wop;
        const int hop = (int)(r % Hop);
        const int b = (int)(r / Hop);
        const int boff = f16e5_off(c8 * 8);
        uint4 rh = make_uint4(0, 0, 0, 0);
        uint2 ra = make_uint2(0, 0), rb = make_uint2(0, 0);
        if (wop > 0 && hop < Ho) {
            const int h0 = hop * 2, w0 = (wop - 1) * 2;
            float best[8];
            unsigned short bh[8];
            uint8_t ba[8], bb[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { best[e] = -3.4e38f; bh[e] = 0; ba[e] = 0; bb[e] = 0; }
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const long long row = ((long long)b * Hp + h0 + dy) * Wp + (w0 + dx + 1);
                    const uint4 vh = *reinterpret_cast<const uint4*>(ih + row * c_pad + c8 * 8);
                    const uint2 va = *reinterpret_cast<const uint2*>(il + row * c_pad * 2 + boff);
                    const uint2 vb = *reinterpret_cast<const uint2*>(il + row * c_pad * 2 + boff + 64);
                    const unsigned short* ph = reinterpret_cast<const unsigned short*>(&vh);
                    const uint8_t* pa = reinterpret_cast<const uint8_t*>(&va);
                    const uint8_t* pb = reinterpret_cast<const uint8_t*>(&vb);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float x = join_f16e5(ph[e], pb[e]);
                        if (x > best[e]) { best[e] = x; bh[e] = ph[e]; ba[e] = pa[e]; bb[e] = pb[e]; }
                    }
                }
            rh = *reinterpret_cast<uint4*>(bh);
            ra = *reinterpret_cast<uint2*>(ba);
            rb = *reinterpret_cast<uint2*>(bb);
        }
        *reinterpret_cast<uint4*>(oh + orow * c_pad + c8 * 8) = rh;
        *reinterpret_cast<uint2*>(ol + orow * c_pad * 2 + boff) = ra;
        *reinterpret_cast<uint2*>(ol + orow * c_pad * 2 + boff + 64) = rb;
    }
}

static inline int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148LL * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace mv3d

using namespace mv3d;

extern "C" __attribute__((visibility("default"))) int mv3d_pack_weights(const float* d_w, int taps, int cin, int cout, int cin_pad, void* d_hi, void* d_lo,
                                 void* stream) {
    MV3D_REQUIRE(d_w && d_hi && taps > 0 && cin > 0 && cout > 0 && cin_pad >= cin);
    const long long total = (long long)cout * taps * cin_pad;
    if (cout >= 64 && cin_pad % 8 == 0 && (long long)taps * cin_pad <= 65535LL * 64 &&
        (reinterpret_cast<uintptr_t>(d_hi) & 15) == 0 && (!d_lo || (reinterpret_cast<uintptr_t>(d_lo) & 15) == 0)) {
        dim3 grid(ceil_div(cout, 64), ceil_div(taps * cin_pad, 64));
        pack_weights_tiled64_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
            d_w, taps, cin, cout, cin_pad, (__nv_bfloat16*)d_hi, (__nv_bfloat16*)d_lo);
    } else if (cout >= 32 && (long long)taps * cin_pad <= 65535LL * 32) {
        dim3 grid(ceil_div(cout, 32), ceil_div(taps * cin_pad, 32));
        pack_weights_tiled_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(
            d_w, taps, cin, cout, cin_pad, (__nv_bfloat16*)d_hi, (__nv_bfloat16*)d_lo);
    } else {
        pack_weights_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
            d_w, taps, cin, cout, cin_pad, (__nv_bfloat16*)d_hi, (__nv_bfloat16*)d_lo);
    }
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

extern "C" __attribute__((visibility("default"))) int mv3d_pad_nhwc(const float* d_in, int B, int H, int W, int C, int c_pad, void* d_hi, void* d_lo,
                             void* stream) {
    MV3D_REQUIRE(d_in && d_hi && B > 0 && H > 0 && W > 0 && C > 0 && c_pad >= C && c_pad % 8 == 0);
    const long long total = (long long)B * (H + 1) * (W + 1) * (c_pad / 8);
    pad_nhwc_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(d_in, B, H, W, C, c_pad,
                                                                             (__nv_bfloat16*)d_hi, (__nv_bfloat16*)d_lo);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

extern "C" __attribute__((visibility("default"))) int mv3d_im2col3x3_pad(const float* d_in, int B, int H, int W, int C, int k_pad,
                                                                          void* d_hi, void* d_lo, void* stream) {
    MV3D_REQUIRE(d_in && d_hi && B > 0 && H > 0 && W > 0 && C > 0 && k_pad >= 9 * C && k_pad % 8 == 0);
    const long long total = (long long)B * (H + 1) * (W + 1) * (k_pad / 8);
    im2col3x3_pad_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(d_in, B, H, W, C, k_pad,
                                                                                  (__nv_bfloat16*)d_hi, (__nv_bfloat16*)d_lo);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

extern "C" __attribute__((visibility("default"))) int mv3d_unpad_nhwc(const void* d_hi, const void* d_lo, int B, int H, int W, int C, int c_pad,
                               float* d_out, void* stream) {
    MV3D_REQUIRE(d_hi && d_out && B > 0 && H > 0 && W > 0 && C > 0 && c_pad >= C);
    const long long total = (long long)B * H * W * C;
    unpad_nhwc_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)d_hi, (const __nv_bfloat16*)d_lo, B, H, W, C, c_pad, d_out);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

extern "C" __attribute__((visibility("default"))) int mv3d_maxpool2x2_pad(const void* d_in_hi, const void* d_in_lo, int B, int H, int W, int c_pad,
                                   void* d_out_hi, void* d_out_lo, void* stream) {
    MV3D_REQUIRE(d_in_hi && d_out_hi && B > 0 && H > 1 && W > 1 && c_pad % 8 == 0);
    const long long total = (long long)B * (H / 2 + 1) * (W / 2 + 1) * (c_pad / 8);
    maxpool2x2_pad_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)d_in_hi, (const __nv_bfloat16*)d_in_lo, B, H, W, c_pad, (__nv_bfloat16*)d_out_hi,
        (__nv_bfloat16*)d_out_lo);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

// ---- MV3D_FMT_F16E5 variants (fmt = MV3D_FMT_BF16X2 forwards to the functions above) ----
extern "C" __attribute__((visibility("default"))) int mv3d_pack_weights_fmt(const float* d_w, int taps, int cin, int cout, int cin_pad, void* d_hi,
                                                                           void* d_lo, int fmt, void* stream) {
    if (fmt == MV3D_FMT_BF16X2) return mv3d_pack_weights(d_w, taps, cin, cout, cin_pad, d_hi, d_lo, stream);
    MV3D_REQUIRE(fmt == MV3D_FMT_F16E5 && d_w && d_hi && d_lo && taps > 0 && cin > 0 && cout > 0 && cin_pad >= cin && cin_pad % 64 == 0);
    MV3D_REQUIRE((long long)taps * cin_pad <= 65535LL * 32);
    dim3 grid(ceil_div(cout, 32), ceil_div(taps * cin_pad, 32));
    pack_weights_f16e5_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(d_w, taps, cin, cout, cin_pad,
                                                                              (unsigned short*)d_hi, (uint8_t*)d_lo);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

extern "C" __attribute__((visibility("default"))) int mv3d_pad_nhwc_fmt(const float* d_in, int B, int H, int W, int C, int c_pad, void* d_hi,
                                                                       void* d_lo, int fmt, void* stream) {
    if (fmt == MV3D_FMT_BF16X2) return mv3d_pad_nhwc(d_in, B, H, W, C, c_pad, d_hi, d_lo, stream);
    MV3D_REQUIRE(fmt == MV3D_FMT_F16E5 && d_in && d_hi && d_lo && B > 0 && H > 0 && W > 0 && C > 0 && c_pad >= C && c_pad % 64 == 0);
    const long long total = (long long)B * (H + 1) * (W + 1) * (c_pad / 8);
    pad_nhwc_f16e5_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(d_in, B, H, W, C, c_pad,
                                                                                   (unsigned short*)d_hi, (uint8_t*)d_lo);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

extern "C" __attribute__((visibility("default"))) int mv3d_unpad_nhwc_fmt(const void* d_hi, const void* d_lo, int B, int H, int W, int C, int c_pad,
                                                                         float* d_out, int fmt, void* stream) {
    if (fmt == MV3D_FMT_BF16X2) return mv3d_unpad_nhwc(d_hi, d_lo, B, H, W, C, c_pad, d_out, stream);
    MV3D_REQUIRE(fmt == MV3D_FMT_F16E5 && d_hi && d_lo && d_out && B > 0 && H > 0 && W > 0 && C > 0 && c_pad >= C && c_pad % 64 == 0);
    const long long total = (long long)B * H * W * C;
    unpad_nhwc_f16e5_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const unsigned short*)d_hi, (const uint8_t*)d_lo, B, H, W, C, c_pad, d_out);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

extern "C" __attribute__((visibility("default"))) int mv3d_maxpool2x2_pad_fmt(const void* d_in_hi, const void* d_in_lo, int B, int H, int W, int c_pad,
                                                                             void* d_out_hi, void* d_out_lo, int fmt, void* stream) {
    if (fmt == MV3D_FMT_BF16X2) return mv3d_maxpool2x2_pad(d_in_hi, d_in_lo, B, H, W, c_pad, d_out_hi, d_out_lo, stream);
    MV3D_REQUIRE(fmt == MV3D_FMT_F16E5 && d_in_hi && d_in_lo && d_out_hi && d_out_lo && B > 0 && H > 1 && W > 1 && c_pad % 64 == 0);
    const long long total = (long long)B * (H / 2 + 1) * (W / 2 + 1) * (c_pad / 8);
    maxpool2x2_pad_f16e5_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const unsigned short*)d_in_hi, (const uint8_t*)d_in_lo, B, H, W, c_pad, (unsigned short*)d_out_hi,
        (uint8_t*)d_out_lo);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

extern "C" __attribute__((visibility("default"))) int mv3d_softmax_pairs(const float* d_in, int rows, int ld_in, int n_pairs, float* d_out, int ld_out,
                                  void* stream) {
    MV3D_REQUIRE(d_in && d_out && rows > 0 && n_pairs > 0 && ld_in >= 2 * n_pairs && ld_out >= 2 * n_pairs);
    softmax_pairs_kernel<<<grid_for((long long)rows * n_pairs, 256), 256, 0, (cudaStream_t)stream>>>(
        d_in, rows, ld_in, n_pairs, d_out, ld_out);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

extern "C" __attribute__((visibility("default"))) int mv3d_bias_act(const float* d_acc, int M, int N, int ld_acc, const float* d_bias, int relu,
                             void* d_out_hi, void* d_out_lo, int ld_out, float* d_out_f32, int ld_f32, void* stream) {
    MV3D_REQUIRE(d_acc && M > 0 && N > 0 && (d_out_hi || d_out_f32));
    bias_act_kernel<<<grid_for((long long)M * N, 256), 256, 0, (cudaStream_t)stream>>>(
        d_acc, M, N, ld_acc, d_bias, relu, (__nv_bfloat16*)d_out_hi, (__nv_bfloat16*)d_out_lo, ld_out, d_out_f32,
        ld_f32);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

namespace mv3d {
int launch_small_cin_mma(const float* d_in, int B, int H, int W, int C, const float* d_w, const float* d_bias, int relu,
                         void* d_out_hi, void* d_out_lo, int fmt, cudaStream_t st);   // first_layer_tcgen05.cu
}
/* Direct 3x3 SAME conv (+bias, +ReLU) of a dense float32 (B,H,W,C) input with C <= 4 into the PAD layout in `fmt`
 * (Network.conv on the RGB / front-view image, lib/networks/network.py:108-132 with MV3D_test.py:51): replaces
 * mv3d_im2col3x3_pad + the K = 32 GEMM in inference.  d_w: HWIO (3,3,C,Cout) float32; Cout % 8 == 0, c_pad >= Cout. */
extern "C" __attribute__((visibility("default"))) int mv3d_conv3x3_small_cin(
    const float* d_in, int B, int H, int W, int C, const float* d_w, const float* d_bias, int Cout, int relu,
    void* d_out_hi, void* d_out_lo, int c_pad, int fmt, void* stream) {
    MV3D_REQUIRE(d_in && d_w && d_out_hi && B > 0 && H > 0 && W > 0 && C > 0 && C <= 4 && Cout > 0 && Cout % 8 == 0);
    MV3D_REQUIRE(c_pad >= Cout && c_pad % 8 == 0);
    MV3D_REQUIRE(fmt == MV3D_FMT_BF16X2 || (fmt == MV3D_FMT_F16E5 && d_out_lo && c_pad % 64 == 0));
    cudaStream_t st = (cudaStream_t)stream;
    static int mma_mode = -1;   // MV3D_SMALL_CIN_MMA=0: always the direct fp32 kernel (A/B comparisons)
    if (mma_mode < 0) { const char* e = getenv("MV3D_SMALL_CIN_MMA"); mma_mode = e ? atoi(e) : 1; }
    if (mma_mode && C <= 3 && Cout == 64 && c_pad == 64)
        return mv3d::launch_small_cin_mma(d_in, B, H, W, C, d_w, d_bias, relu, d_out_hi, d_out_lo, fmt, st);
    const size_t smem = sizeof(float) * ((size_t)9 * C * Cout + Cout);
    MV3D_REQUIRE(smem <= 48 * 1024);
    const long long total = (long long)B * (H + 1) * ((W + 1 + kSmallPx - 1) / kSmallPx) * (c_pad / 8);
    const int grid = grid_for(total, 256);
#define MV3D_SMALL_CIN(FMT, CC) \
    conv3x3_small_cin_kernel<FMT, CC><<<grid, 256, smem, st>>>(d_in, B, H, W, d_w, d_bias, Cout, relu, d_out_hi, d_out_lo, c_pad)
    if (fmt == MV3D_FMT_F16E5) {
        if (C == 1) MV3D_SMALL_CIN(MV3D_FMT_F16E5, 1); else if (C == 2) MV3D_SMALL_CIN(MV3D_FMT_F16E5, 2);
        else if (C == 3) MV3D_SMALL_CIN(MV3D_FMT_F16E5, 3); else MV3D_SMALL_CIN(MV3D_FMT_F16E5, 4);
    } else {
        if (C == 1) MV3D_SMALL_CIN(MV3D_FMT_BF16X2, 1); else if (C == 2) MV3D_SMALL_CIN(MV3D_FMT_BF16X2, 2);
        else if (C == 3) MV3D_SMALL_CIN(MV3D_FMT_BF16X2, 3); else MV3D_SMALL_CIN(MV3D_FMT_BF16X2, 4);
    }
#undef MV3D_SMALL_CIN
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}
