// Training-side elementwise / reduction kernels around the tcgen05 GEMMs: max-pool backward on the PAD layout,
// bias gradient, backward-data weight packing, dense->PAD with ReLU gate, dropout, the four MV3D losses with their
// gradients (lib/fast_rcnn/train_mv.py:67-84,94-139) and Adam (train_mv.py:144-146, TF defaults).  All HBM-bound.
#include "common.cuh"

namespace mv3d {

static inline int grid_for_t(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148LL * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

__device__ __forceinline__ float pair_val(__nv_bfloat16 h, __nv_bfloat16 l) {
    return __bfloat162float(h) + __bfloat162float(l);
}

// ------------------------------------------------------------------------------------------------
// Backward of Network.max_pool(2,2,2,2,'VALID') (network.py:181-188) fused with the ReLU gate of the conv that
// produced its input.  x = pre-pool activation (PAD, H x W), g = gradient w.r.t. the pooled map (PAD, H/2 x W/2).
// The gradient goes to the FIRST maximum of each window in (dy,dx) order -- the same element the forward kernel
// (maxpool2x2_pad_kernel, strict `>`) selected -- and only if that maximum is > 0 (ReLU').  Output: PAD H x W.
// ------------------------------------------------------------------------------------------------
// One thread per 2x2 WINDOW and 8 channels: the four inputs and the pooled gradient are read once (the per-output form
// re-read every window four times through L2), the four outputs written as 16-byte stores.  Threads of the extra row /
// column groups write the zeros of the halo (PAD column 0, row H) and of an odd trailing row / column.
__global__ void maxpool2x2_bwd_pad_kernel(const __nv_bfloat16* __restrict__ xh, const __nv_bfloat16* __restrict__ xl,
                                          const __nv_bfloat16* __restrict__ gh, const __nv_bfloat16* __restrict__ gl,
                                          int B, int H, int W, int c_pad, __nv_bfloat16* __restrict__ oh,
                                          __nv_bfloat16* __restrict__ ol) {
    const int Ho = H / 2, Wo = W / 2;
    const int Hp = H + 1, Wp = W + 1, Hop = Ho + 1, Wop = Wo + 1;
    const int cv = c_pad / 8;
    const int ncg = Wo + 1 + (W & 1);    // column groups: {0}, {1,2}, {3,4}, ..., and {W} when W is odd
    const int nrg = Ho + 1;              // row groups: {0,1}, {2,3}, ..., and the tail rows 2*Ho .. H (one or two)
    const long long total = (long long)B * nrg * ncg * cv;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cv);
        long long r = i / cv;
        const int cg = (int)(r % ncg);
        r /= ncg;
        const int rg = (int)(r % nrg);
        const int b = (int)(r / nrg);
        const int row0 = 2 * rg, nrows = rg < Ho ? 2 : Hp - 2 * Ho;
        const int col0 = cg == 0 ? 0 : 2 * cg - 1, ncols = (cg == 0 || cg > Wo) ? 1 : 2;
        if (rg < Ho && cg >= 1 && cg <= Wo) {
            float v[4][8];
            long long j[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                j[k] = ((((long long)b * Hp + row0 + (k >> 1)) * Wp + (col0 + (k & 1))) * c_pad) + c8 * 8;
                const uint4 vh = *reinterpret_cast<const uint4*>(xh + j[k]);
                const uint4 vl = xl ? *reinterpret_cast<const uint4*>(xl + j[k]) : zero;
                const __nv_bfloat16* a = reinterpret_cast<const __nv_bfloat16*>(&vh);
                const __nv_bfloat16* c = reinterpret_cast<const __nv_bfloat16*>(&vl);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[k][e] = pair_val(a[e], c[e]);
            }
            const long long gj = ((((long long)b * Hop + rg) * Wop + cg) * c_pad) + c8 * 8;
            const uint4 g4h = *reinterpret_cast<const uint4*>(gh + gj);
            const uint4 g4l = gl ? *reinterpret_cast<const uint4*>(gl + gj) : zero;
            const uint32_t* pgh = reinterpret_cast<const uint32_t*>(&g4h);
            const uint32_t* pgl = reinterpret_cast<const uint32_t*>(&g4l);
            uint32_t outh[4][4], outl[4][4];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                // the FIRST maximum in (dy,dx) scan order takes the gradient (forward: strict `>`), and only if it is > 0
                int win = 0;
                float best = v[0][e];
#pragma unroll
                for (int k = 1; k < 4; ++k)
                    if (v[k][e] > best) { best = v[k][e]; win = k; }
                if (!(best > 0.f)) win = -1;
                const uint32_t sel = (e & 1) ? 0xffff0000u : 0x0000ffffu;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if ((e & 1) == 0) { outh[k][e >> 1] = 0u; outl[k][e >> 1] = 0u; }
                    if (k == win) { outh[k][e >> 1] |= pgh[e >> 1] & sel; outl[k][e >> 1] |= pgl[e >> 1] & sel; }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                *reinterpret_cast<uint4*>(oh + j[k]) = make_uint4(outh[k][0], outh[k][1], outh[k][2], outh[k][3]);
                if (ol) *reinterpret_cast<uint4*>(ol + j[k]) = make_uint4(outl[k][0], outl[k][1], outl[k][2], outl[k][3]);
            }
        } else {
            for (int dy = 0; dy < nrows; ++dy)
                for (int dx = 0; dx < ncols; ++dx) {
                    const long long jz = ((((long long)b * Hp + row0 + dy) * Wp + (col0 + dx)) * c_pad) + c8 * 8;
                    *reinterpret_cast<uint4*>(oh + jz) = zero;
                    if (ol) *reinterpret_cast<uint4*>(ol + jz) = zero;
                }
        }
    }
}

// db[n] += sum_rows (g_hi + g_lo)[row, n].  Block = 64 channel pairs x 4 row lanes; rows strided over the grid.
__global__ void bias_grad_kernel(const __nv_bfloat16* __restrict__ gh, const __nv_bfloat16* __restrict__ gl,
                                 long long rows, int ld, int n, float* __restrict__ db) {
    __shared__ float red[4][128];
    const int cx = threadIdx.x & 63, ry = threadIdx.x >> 6;
    for (int c0 = 0; c0 < n; c0 += 128) {
        const int c = c0 + cx * 2;
        float s0 = 0.f, s1 = 0.f;
        if (c < ld) {
            for (long long r = (long long)blockIdx.x * 4 + ry; r < rows; r += (long long)gridDim.x * 4) {
                const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(gh + r * ld + c);
                s0 += __bfloat162float(a.x);
                s1 += __bfloat162float(a.y);
                if (gl) {
                    const __nv_bfloat162 l2 = *reinterpret_cast<const __nv_bfloat162*>(gl + r * ld + c);
                    s0 += __bfloat162float(l2.x);
                    s1 += __bfloat162float(l2.y);
                }
            }
        }
        red[ry][cx * 2] = s0;
        red[ry][cx * 2 + 1] = s1;
        __syncthreads();
        if (threadIdx.x < 128) {
            const int cc = c0 + threadIdx.x;
            if (cc < n) atomicAdd(db + cc, red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x]);
        }
        __syncthreads();
    }
}

// Vectorised form for row pitches that are multiples of 8 and <= 2048: a thread owns 8 consecutive channels (16-byte
// loads) of every (blockDim.x / (ld/8))-th row; partial sums meet in shared memory, one atomicAdd per channel and CTA.
__global__ void bias_grad_vec_kernel(const __nv_bfloat16* __restrict__ gh, const __nv_bfloat16* __restrict__ gl,
                                     long long rows, int ld, int n, float* __restrict__ db) {
    extern __shared__ float sm[];   // [row_lanes][ld]
    const int cv = ld / 8;
    const int row_lanes = blockDim.x / cv;
    const int v = threadIdx.x % cv, rl = threadIdx.x / cv;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (rl < row_lanes) {
        for (long long r = (long long)blockIdx.x * row_lanes + rl; r < rows; r += (long long)gridDim.x * row_lanes) {
            const uint4 a = *reinterpret_cast<const uint4*>(gh + r * ld + v * 8);
            const __nv_bfloat16* pa = reinterpret_cast<const __nv_bfloat16*>(&a);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += __bfloat162float(pa[e]);
            if (gl) {
                const uint4 b = *reinterpret_cast<const uint4*>(gl + r * ld + v * 8);
                const __nv_bfloat16* pb = reinterpret_cast<const __nv_bfloat16*>(&b);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] += __bfloat162float(pb[e]);
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) sm[rl * ld + v * 8 + e] = acc[e];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < row_lanes; ++k) s += sm[k * ld + c];
        atomicAdd(db + c, s);
    }
}

// Backward-data weight: HWIO fp32 (taps, cin, cout) -> bf16 hi/lo (cin, taps*cout_pad) with the taps FLIPPED
// (tap' = taps-1-tap): dX[p, c] = sum_{t'} sum_n G[p + shift_{t'}, n] * W[taps-1-t', c, n].
__global__ void pack_weights_dgrad_kernel(const float* __restrict__ w, int taps, int cin, int cout, int cout_pad,
                                          __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const long long total = (long long)cin * taps * cout_pad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i % cout_pad);
        const long long r = i / cout_pad;
        const int tp = (int)(r % taps);
        const int c = (int)(r / taps);
        float x = 0.f;
        if (n < cout) x = w[((long long)(taps - 1 - tp) * cin + c) * cout + n];
        __nv_bfloat16 h, l;
        split_bf16(x, h, l);
        hi[i] = h;
        if (lo) lo[i] = l;
    }
}

// Same packing, eight consecutive output channels per thread (two float4 loads, one 16-byte store per plane); used when
// cout % 4 == 0 so that the loads are aligned.
__global__ void pack_weights_dgrad_vec_kernel(const float* __restrict__ w, int taps, int cin, int cout, int cout_pad,
                                              __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const int nv = cout_pad / 8;
    const long long total = (long long)cin * taps * nv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int n8 = (int)(i % nv) * 8;
        const long long r = i / nv;
        const int tp = (int)(r % taps);
        const int c = (int)(r / taps);
        const float* src = w + ((long long)(taps - 1 - tp) * cin + c) * cout + n8;
        float x[8];
        if (n8 + 8 <= cout) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src));
            const float4 b = __ldg(reinterpret_cast<const float4*>(src + 4));
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = (n8 + e < cout) ? src[e] : 0.f;
        }
        uint32_t ph[4], pl[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(x[2 * e], h0, l0);
            split_bf16(x[2 * e + 1], h1, l1);
            ph[e] = uint32_t(__bfloat16_as_ushort(h0)) | (uint32_t(__bfloat16_as_ushort(h1)) << 16);
            pl[e] = uint32_t(__bfloat16_as_ushort(l0)) | (uint32_t(__bfloat16_as_ushort(l1)) << 16);
        }
        *reinterpret_cast<uint4*>(hi + i * 8) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        if (lo) *reinterpret_cast<uint4*>(lo + i * 8) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
}

// dense fp32 (B,H,W,C) gradient -> PAD bf16 hi/lo gated by (mask_hi > 0) (mask = forward activation, PAD, same C pad).
__global__ void pad_nhwc_masked_kernel(const float* __restrict__ in, int B, int H, int W, int C, int c_pad,
                                       const __nv_bfloat16* __restrict__ mask, __nv_bfloat16* __restrict__ hi,
                                       __nv_bfloat16* __restrict__ lo) {
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
            float x = (inside && c < C) ? src[c] : 0.f;
            if (mask && !(__bfloat162float(mask[i * 8 + e]) > 0.f)) x = 0.f;
            split_bf16(x, vh[e], vl[e]);
        }
        *reinterpret_cast<uint4*>(hi + i * 8) = *reinterpret_cast<uint4*>(vh);
        if (lo) *reinterpret_cast<uint4*>(lo + i * 8) = *reinterpret_cast<uint4*>(vl);
    }
}

// Counter-based RNG (splitmix64 finaliser over (seed, index)) -> uniform [0,1).
__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned long long idx) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// Network.dropout (network.py:407-409 -> tf.nn.dropout): y = x / keep_prob where u < keep_prob, else 0; in place.
__global__ void dropout_kernel(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long rows, int n,
                               int ld, float keep_prob, unsigned long long seed) {
    const long long total = rows * n;
    const float inv = 1.0f / keep_prob;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / n;
        const int c = (int)(i % n);
        const long long j = r * ld + c;
        float x = __bfloat162float(hi[j]) + (lo ? __bfloat162float(lo[j]) : 0.f);
        x = (uniform01(seed, (unsigned long long)i) < keep_prob) ? x * inv : 0.f;
        __nv_bfloat16 h, l;
        split_bf16(x, h, l);
        hi[j] = h;
        if (lo) lo[j] = l;
    }
}

__device__ __forceinline__ void smooth_l1(float d, float sigma2, float& loss, float& grad) {
    // train_mv.py:67-84: 0.5*sigma2*d^2 if |d| < 1/sigma2 else |d| - 0.5/sigma2
    const float a = fabsf(d);
    if (a < 1.0f / sigma2) {
        loss = d * d * (0.5f * sigma2);
        grad = sigma2 * d;
    } else {
        loss = a - 0.5f / sigma2;
        grad = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    }
}

// RPN losses and their gradients (train_mv.py:94-119).  One thread per PAD pixel of the RPN map (halo -> zeros).
//   cls_score (B,Hf,Wf,2A), bbox_pred (B,Hf,Wf,6A) dense fp32; labels (B,Hf,Wf,A) float32 {-1,0,1};
//   targets (B,Hf*Wf*A,6); counts (B,2) int32 = {#label != -1, #label == 1} per frame.
//   Gradient written as PAD bf16 hi/lo (B,Hf+1,Wf+1,c_pad): channels [0,2A) = d cls_score, [2A, 8A) = d bbox_pred.
//   loss_out[0] += rpn_cross_entropy, loss_out[1] += rpn_loss_box (both already averaged over the B frames).
__global__ void rpn_loss_kernel(const float* __restrict__ cls, const float* __restrict__ bbox,
                                const float* __restrict__ labels, const float* __restrict__ targets,
                                const int* __restrict__ counts, int B, int Hf, int Wf, int A, int c_pad,
                                float sigma2, __nv_bfloat16* __restrict__ gh, __nv_bfloat16* __restrict__ gl,
                                float* __restrict__ loss_out) {
    const int Hp = Hf + 1, Wp = Wf + 1;
    const long long total = (long long)B * Hp * Wp;
    float l_cls = 0.f, l_box = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int wp = (int)(i % Wp);
        long long r = i / Wp;
        const int hp = (int)(r % Hp);
        const int b = (int)(r / Hp);
        __nv_bfloat16* oh = gh + i * c_pad;
        __nv_bfloat16* ol = gl ? gl + i * c_pad : nullptr;
        const bool inside = wp > 0 && hp < Hf;
        if (!inside) {
            for (int c = 0; c < c_pad; ++c) {
                oh[c] = __float2bfloat16_rn(0.f);
                if (ol) ol[c] = __float2bfloat16_rn(0.f);
            }
            continue;
        }
        const long long pix = ((long long)b * Hf + hp) * Wf + (wp - 1);
        const float n_valid = (float)counts[b * 2], n_pos = (float)counts[b * 2 + 1];
        const float w_cls = n_valid > 0.f ? 1.0f / (n_valid * B) : 0.f;
        const float w_box = n_pos > 0.f ? 1.0f / (n_pos * B) : 0.f;
        for (int a = 0; a < A; ++a) {
            const float lab = labels[pix * A + a];
            float g0 = 0.f, g1 = 0.f;
            if (lab != -1.0f) {
                const float s0 = cls[pix * 2 * A + 2 * a], s1 = cls[pix * 2 * A + 2 * a + 1];
                const float m = fmaxf(s0, s1);
                const float e0 = expf(s0 - m), e1 = expf(s1 - m);
                const float lse = m + logf(e0 + e1);
                const float p0 = e0 / (e0 + e1), p1 = e1 / (e0 + e1);
                l_cls += (lse - (lab == 1.0f ? s1 : s0)) * w_cls;
                g0 = (p0 - (lab == 1.0f ? 0.f : 1.f)) * w_cls;
                g1 = (p1 - (lab == 1.0f ? 1.f : 0.f)) * w_cls;
            }
            __nv_bfloat16 h, l;
            split_bf16(g0, h, l); oh[2 * a] = h; if (ol) ol[2 * a] = l;
            split_bf16(g1, h, l); oh[2 * a + 1] = h; if (ol) ol[2 * a + 1] = l;
            for (int k = 0; k < 6; ++k) {
                float g = 0.f;
                if (lab == 1.0f) {
                    const float d = bbox[pix * 6 * A + 6 * a + k] - targets[(pix * A + a) * 6 + k];
                    float ls;
                    smooth_l1(d, sigma2, ls, g);
                    l_box += ls * w_box;
                    g *= w_box;
                }
                split_bf16(g, h, l);
                oh[2 * A + 6 * a + k] = h;
                if (ol) ol[2 * A + 6 * a + k] = l;
            }
        }
        for (int c = 8 * A; c < c_pad; ++c) {
            oh[c] = __float2bfloat16_rn(0.f);
            if (ol) ol[c] = __float2bfloat16_rn(0.f);
        }
    }
    // block reduction of the two loss terms
    __shared__ float sred[2][32];
    for (int o = 16; o > 0; o >>= 1) {
        l_cls += __shfl_xor_sync(0xffffffffu, l_cls, o);
        l_box += __shfl_xor_sync(0xffffffffu, l_box, o);
    }
    if ((threadIdx.x & 31) == 0) { sred[0][threadIdx.x >> 5] = l_cls; sred[1][threadIdx.x >> 5] = l_box; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, c = 0.f;
        for (int k = 0; k < (blockDim.x + 31) / 32; ++k) { a += sred[0][k]; c += sred[1][k]; }
        atomicAdd(loss_out + 0, a);
        atomicAdd(loss_out + 1, c);
    }
}

// R-CNN losses and gradients (train_mv.py:121-133).  One thread per sampled roi row.
//   cls_score (R,2) / bbox_pred (R,nb) fp32 with pitches; labels (R) int32; targets (R,nb); roi_batch (R) float = rois[:,0];
//   frame_counts (B) int32 = rois per frame.  Gradient -> bf16 hi/lo (R, c_pad): [0,2) d cls_score, [2,2+nb) d bbox_pred.
__global__ void rcnn_loss_kernel(const float* __restrict__ cls, int ld_cls, const float* __restrict__ bbox, int ld_bbox,
                                 const int* __restrict__ labels, const float* __restrict__ targets, int nb,
                                 const float* __restrict__ rois, const int* __restrict__ frame_counts, int B, int R,
                                 int c_pad, float sigma2, __nv_bfloat16* __restrict__ gh,
                                 __nv_bfloat16* __restrict__ gl, float* __restrict__ loss_out) {
    float l_cls = 0.f, l_box = 0.f;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
        const int b = (int)rois[r * 5];
        const float nf = (float)frame_counts[b];
        const float wgt = nf > 0.f ? 1.0f / (nf * B) : 0.f;
        const int lab = labels[r];
        const float s0 = cls[(long long)r * ld_cls], s1 = cls[(long long)r * ld_cls + 1];
        const float m = fmaxf(s0, s1);
        const float e0 = expf(s0 - m), e1 = expf(s1 - m);
        const float lse = m + logf(e0 + e1);
        l_cls += (lse - (lab == 1 ? s1 : s0)) * wgt;
        __nv_bfloat16* oh = gh + (long long)r * c_pad;
        __nv_bfloat16* ol = gl ? gl + (long long)r * c_pad : nullptr;
        __nv_bfloat16 h, l;
        split_bf16((e0 / (e0 + e1) - (lab == 1 ? 0.f : 1.f)) * wgt, h, l); oh[0] = h; if (ol) ol[0] = l;
        split_bf16((e1 / (e0 + e1) - (lab == 1 ? 1.f : 0.f)) * wgt, h, l); oh[1] = h; if (ol) ol[1] = l;
        for (int k = 0; k < nb; ++k) {
            const float d = bbox[(long long)r * ld_bbox + k] - targets[(long long)r * nb + k];
            float ls, g;
            smooth_l1(d, sigma2, ls, g);
            l_box += ls * wgt;
            split_bf16(g * wgt, h, l);
            oh[2 + k] = h;
            if (ol) ol[2 + k] = l;
        }
        for (int c = 2 + nb; c < c_pad; ++c) {
            oh[c] = __float2bfloat16_rn(0.f);
            if (ol) ol[c] = __float2bfloat16_rn(0.f);
        }
    }
    __shared__ float sred[2][32];
    for (int o = 16; o > 0; o >>= 1) {
        l_cls += __shfl_xor_sync(0xffffffffu, l_cls, o);
        l_box += __shfl_xor_sync(0xffffffffu, l_box, o);
    }
    if ((threadIdx.x & 31) == 0) { sred[0][threadIdx.x >> 5] = l_cls; sred[1][threadIdx.x >> 5] = l_box; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, c = 0.f;
        for (int k = 0; k < (blockDim.x + 31) / 32; ++k) { a += sred[0][k]; c += sred[1][k]; }
        atomicAdd(loss_out + 0, a);
        atomicAdd(loss_out + 1, c);
    }
}

// tf.train.AdamOptimizer (TF 1.0 defaults beta1=0.9, beta2=0.999, epsilon=1e-8; train_mv.py:144-146):
//   lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; theta -= lr_t*m/(sqrt(v)+eps).
// grad_scale folds the 1/world_size of the data-parallel gradient all-reduce into the update.
__device__ __forceinline__ void adam_one(float& th, float g, float& mi, float& vi, float lr_t, float b1, float b2,
                                         float eps, float grad_scale) {
    g *= grad_scale;
    mi = b1 * mi + (1.0f - b1) * g;
    vi = b2 * vi + (1.0f - b2) * g * g;
    th -= lr_t * mi / (sqrtf(vi) + eps);
}

__global__ void adam_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr_t, float b1, float b2, float eps,
                            float grad_scale) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        adam_one(theta[i], grad[i], m[i], v[i], lr_t, b1, b2, eps, grad_scale);
}

// 16-byte form (all four arrays 16-byte aligned): 28 bytes per parameter in 7 vector accesses per 4 parameters; same
// per-element arithmetic as adam_kernel.  The n % 4 tail is handled by the scalar kernel.
__global__ void adam_vec4_kernel(float4* __restrict__ theta, const float4* __restrict__ grad, float4* __restrict__ m,
                                 float4* __restrict__ v, long long n4, float lr_t, float b1, float b2, float eps,
                                 float grad_scale) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        float4 t = theta[i], mm = m[i], vv = v[i];
        const float4 g = __ldcs(grad + i);      // read once
        adam_one(t.x, g.x, mm.x, vv.x, lr_t, b1, b2, eps, grad_scale);
        adam_one(t.y, g.y, mm.y, vv.y, lr_t, b1, b2, eps, grad_scale);
        adam_one(t.z, g.z, mm.z, vv.z, lr_t, b1, b2, eps, grad_scale);
        adam_one(t.w, g.w, mm.w, vv.w, lr_t, b1, b2, eps, grad_scale);
        theta[i] = t; m[i] = mm; v[i] = vv;
    }
}

}  // namespace mv3d

using namespace mv3d;
#define MV3D_API extern "C" __attribute__((visibility("default")))

MV3D_API int mv3d_maxpool2x2_bwd_pad(const void* d_x_hi, const void* d_x_lo, const void* d_g_hi, const void* d_g_lo,
                                     int B, int H, int W, int c_pad, void* d_out_hi, void* d_out_lo, void* stream) {
    MV3D_REQUIRE(d_x_hi && d_g_hi && d_out_hi && B > 0 && H > 1 && W > 1 && c_pad % 8 == 0);
    const long long total = (long long)B * (H / 2 + 1) * (W / 2 + 1 + (W & 1)) * (c_pad / 8);
    maxpool2x2_bwd_pad_kernel<<<grid_for_t(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)d_x_hi, (const __nv_bfloat16*)d_x_lo, (const __nv_bfloat16*)d_g_hi,
        (const __nv_bfloat16*)d_g_lo, B, H, W, c_pad, (__nv_bfloat16*)d_out_hi, (__nv_bfloat16*)d_out_lo);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

MV3D_API int mv3d_bias_grad(const void* d_g_hi, const void* d_g_lo, long long rows, int ld, int n, float* d_db,
                            void* stream) {
    MV3D_REQUIRE(d_g_hi && d_db && rows > 0 && n > 0 && ld >= n && ld % 2 == 0);
    if (ld % 8 == 0 && ld / 8 <= 256 && ((uintptr_t)d_g_hi & 15) == 0 && (!d_g_lo || ((uintptr_t)d_g_lo & 15) == 0)) {
        const int cv = ld / 8;
        const int threads = 256 / cv * cv;
        const int row_lanes = threads / cv;
        long long blocks = (rows + row_lanes * 16 - 1) / (row_lanes * 16);
        if (blocks > 148 * 4) blocks = 148 * 4;
        if (blocks < 1) blocks = 1;
        bias_grad_vec_kernel<<<(int)blocks, threads, sizeof(float) * (size_t)row_lanes * ld, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)d_g_hi, (const __nv_bfloat16*)d_g_lo, rows, ld, n, d_db);
        MV3D_CHECK_LAUNCH();
        return MV3D_OK;
    }
    long long blocks = (rows + 63) / 64;
    if (blocks > 148 * 8) blocks = 148 * 8;
    bias_grad_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)d_g_hi,
                                                                    (const __nv_bfloat16*)d_g_lo, rows, ld, n, d_db);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

MV3D_API int mv3d_pack_weights_dgrad(const float* d_w, int taps, int cin, int cout, int cout_pad, void* d_hi,
                                     void* d_lo, void* stream) {
    MV3D_REQUIRE(d_w && d_hi && taps > 0 && cin > 0 && cout > 0 && cout_pad >= cout);
    const long long total = (long long)cin * taps * cout_pad;
    if (cout % 4 == 0 && cout_pad % 8 == 0 && (reinterpret_cast<uintptr_t>(d_w) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(d_hi) & 15) == 0 && (!d_lo || (reinterpret_cast<uintptr_t>(d_lo) & 15) == 0))
        pack_weights_dgrad_vec_kernel<<<grid_for_t(total / 8, 256), 256, 0, (cudaStream_t)stream>>>(
            d_w, taps, cin, cout, cout_pad, (__nv_bfloat16*)d_hi, (__nv_bfloat16*)d_lo);
    else
        pack_weights_dgrad_kernel<<<grid_for_t(total, 256), 256, 0, (cudaStream_t)stream>>>(
            d_w, taps, cin, cout, cout_pad, (__nv_bfloat16*)d_hi, (__nv_bfloat16*)d_lo);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

MV3D_API int mv3d_pad_nhwc_masked(const float* d_in, int B, int H, int W, int C, int c_pad, const void* d_mask_hi,
                                  void* d_hi, void* d_lo, void* stream) {
    MV3D_REQUIRE(d_in && d_hi && B > 0 && H > 0 && W > 0 && C > 0 && c_pad >= C && c_pad % 8 == 0);
    const long long total = (long long)B * (H + 1) * (W + 1) * (c_pad / 8);
    pad_nhwc_masked_kernel<<<grid_for_t(total, 256), 256, 0, (cudaStream_t)stream>>>(
        d_in, B, H, W, C, c_pad, (const __nv_bfloat16*)d_mask_hi, (__nv_bfloat16*)d_hi, (__nv_bfloat16*)d_lo);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

MV3D_API int mv3d_dropout(void* d_hi, void* d_lo, long long rows, int n, int ld, float keep_prob,
                          unsigned long long seed, void* stream) {
    MV3D_REQUIRE(d_hi && rows > 0 && n > 0 && ld >= n && keep_prob > 0.f && keep_prob <= 1.f);
    if (keep_prob >= 1.f) return MV3D_OK;
    dropout_kernel<<<grid_for_t(rows * n, 256), 256, 0, (cudaStream_t)stream>>>(
        (__nv_bfloat16*)d_hi, (__nv_bfloat16*)d_lo, rows, n, ld, keep_prob, seed);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

MV3D_API int mv3d_rpn_loss(const float* d_cls_score, const float* d_bbox_pred, const float* d_labels,
                           const float* d_targets, const int* d_counts, int B, int Hf, int Wf, int A, int c_pad,
                           float sigma, void* d_grad_hi, void* d_grad_lo, float* d_loss, void* stream) {
    MV3D_REQUIRE(d_cls_score && d_bbox_pred && d_labels && d_targets && d_counts && d_grad_hi && d_loss);
    MV3D_REQUIRE(B > 0 && Hf > 0 && Wf > 0 && A > 0 && c_pad >= 8 * A);
    const long long total = (long long)B * (Hf + 1) * (Wf + 1);
    rpn_loss_kernel<<<grid_for_t(total, 128), 128, 0, (cudaStream_t)stream>>>(
        d_cls_score, d_bbox_pred, d_labels, d_targets, d_counts, B, Hf, Wf, A, c_pad, sigma * sigma,
        (__nv_bfloat16*)d_grad_hi, (__nv_bfloat16*)d_grad_lo, d_loss);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

MV3D_API int mv3d_rcnn_loss(const float* d_cls_score, int ld_cls, const float* d_bbox_pred, int ld_bbox,
                            const int* d_labels, const float* d_targets, int n_bbox, const float* d_rois,
                            const int* d_frame_counts, int B, int R, int c_pad, float sigma, void* d_grad_hi,
                            void* d_grad_lo, float* d_loss, void* stream) {
    MV3D_REQUIRE(d_cls_score && d_bbox_pred && d_labels && d_targets && d_rois && d_frame_counts && d_grad_hi && d_loss);
    MV3D_REQUIRE(B > 0 && R > 0 && n_bbox > 0 && c_pad >= 2 + n_bbox);
    rcnn_loss_kernel<<<grid_for_t(R, 64), 64, 0, (cudaStream_t)stream>>>(
        d_cls_score, ld_cls, d_bbox_pred, ld_bbox, d_labels, d_targets, n_bbox, d_rois, d_frame_counts, B, R, c_pad,
        sigma * sigma, (__nv_bfloat16*)d_grad_hi, (__nv_bfloat16*)d_grad_lo, d_loss);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

MV3D_API int mv3d_adam(float* d_theta, const float* d_grad, float* d_m, float* d_v, long long n, float lr, float beta1,
                       float beta2, float eps, int step, float grad_scale, void* stream) {
    MV3D_REQUIRE(d_theta && d_grad && d_m && d_v && n > 0 && step >= 1);
    const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step));
    const bool aligned = (((uintptr_t)d_theta | (uintptr_t)d_grad | (uintptr_t)d_m | (uintptr_t)d_v) & 15) == 0;
    const long long n4 = aligned ? n / 4 : 0;
    if (n4 > 0)
        adam_vec4_kernel<<<grid_for_t(n4, 256), 256, 0, (cudaStream_t)stream>>>(
            (float4*)d_theta, (const float4*)d_grad, (float4*)d_m, (float4*)d_v, n4, (float)lr_t, beta1, beta2, eps, grad_scale);
    if (n - 4 * n4 > 0)
        adam_kernel<<<grid_for_t(n - 4 * n4, 256), 256, 0, (cudaStream_t)stream>>>(
            d_theta + 4 * n4, d_grad + 4 * n4, d_m + 4 * n4, d_v + 4 * n4, n - 4 * n4, (float)lr_t, beta1, beta2, eps, grad_scale);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}
