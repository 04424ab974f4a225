// ROI max pooling, forward (single- and multi-view) and backward.  Replaces the RoiPool / RoiPoolGrad
// TF ops (lib/roi_pooling_layer/roi_pooling_op.cc:74-190,319-452; CUDA roi_pooling_op_gpu.cu.cc:20-215).
//
// The reference runs one thread per OUTPUT ELEMENT with channels innermost in the index but window reads
// strided by C -- and launches once per view.  Here one CTA owns one (view, roi, ph, pw) bin and its threads
// span the channel axis in float4 lanes, so every window pixel is one contiguous 4*C-byte read and every bin
// one contiguous write; all views go in a single launch (grid.y = view).  Feature maps (<= 18 MB) stay in L2,
// so HBM traffic ~= each map once + the pooled outputs.
// Semantics kept bit-exact: round-half-away-from-zero of coord*scale, float32 bin sizes, floor/ceil,
// clamp to [0,H]/[0,W], empty bin -> 0 / argmax -1, strict '>' so the first maximum in (h,w) order wins,
// argmax = (h*W + w)*C + c inside the roi's image.
//
// roi_pool_fused_kernel (the inference path, north-star (iv)): ONE launch pools every view.  One CTA per (roi, view,
// 64-channel slice): thread 0 PROJECTS the 3-D proposal into its view's plane in the kernel (BEV box + clip, 8-corner
// image box, FV box: geom.cuh, the same device functions the proposal layer uses) -- or takes a given rectangle; the
// CTA then STAGES its slice of the roi's whole window (all bins' cells, each read from global memory exactly once,
// 128-bit coalesced, one round of independent loads) in shared memory and every 8-channel lane renders bins from
// shared memory: first maximum in (h, w) order, split into the bf16 hi/lo operand of fc6, written with 16-byte stores.
#include <float.h>

#include "common.cuh"
#include "geom.cuh"

namespace mv3d {

constexpr int kMaxViews = 3;

struct RoiViewDev {
    const float* data;
    const float* rois;
    int H, W;
    float scale;
    float* top;
    int* argmax;
    __nv_bfloat16* top_hi;
    __nv_bfloat16* top_lo;
    int source;        // MV3D_ROI_GIVEN / _BEV / _IMG / _FV (fused kernel)
    float* rois_out;   // optional (R,5): the rectangle that was pooled
    // optional: the feature map in the PAD operand layout the producing conv writes for its other consumers
    // ((B, H+1, W+1, pad_c) 16-bit planes, fmt MV3D_FMT_*); read instead of `data` when set
    const unsigned short* pad_hi;
    const unsigned short* pad_lo;
    int pad_fmt, pad_c;
    int top_fmt;       // rendering of top_hi / top_lo: MV3D_FMT_BF16X2 or MV3D_FMT_F16E5 (fc6 operand, K = bin*C + c)
};
struct RoiProjDev {
    BevGrid bev;
    float M[12];
    const float* dM;
    FvGeom fv;
};
struct RoiViews {
    RoiViewDev v[kMaxViews];
};

struct Bin {
    int batch, hs, he, ws, we;
    bool empty;
};

__device__ __forceinline__ Bin roi_bin(const float* __restrict__ roi, float scale, int H, int W, int PH, int PW, int ph,
                                       int pw) {
    Bin b;
    b.batch = (int)roi[0];
    const int rsw = (int)roundf(roi[1] * scale), rsh = (int)roundf(roi[2] * scale);
    const int rew = (int)roundf(roi[3] * scale), reh = (int)roundf(roi[4] * scale);
    const int rw = max(rew - rsw + 1, 1), rh = max(reh - rsh + 1, 1);
    const float bsh = (float)rh / (float)PH, bsw = (float)rw / (float)PW;
    int hs = (int)floorf((float)ph * bsh), ws = (int)floorf((float)pw * bsw);
    int he = (int)ceilf((float)(ph + 1) * bsh), we = (int)ceilf((float)(pw + 1) * bsw);
    b.hs = min(max(hs + rsh, 0), H); b.he = min(max(he + rsh, 0), H);
    b.ws = min(max(ws + rsw, 0), W); b.we = min(max(we + rsw, 0), W);
    b.empty = (b.he <= b.hs) || (b.we <= b.ws);
    return b;
}

// grid: (R*PH*PW, n_views); block: threads over channel vectors.
template <int VEC>
__global__ void roi_pool_fwd_kernel(RoiViews views, int R, const int* __restrict__ d_num_valid, int C, int PH, int PW) {
    const RoiViewDev& V = views.v[blockIdx.y];
    const int bin = blockIdx.x;
    const int pw = bin % PW, ph = (bin / PW) % PH, n = bin / (PW * PH);
    const size_t out_base = (size_t)bin * C;
    const bool valid = d_num_valid ? (n < *d_num_valid) : true;
    Bin b;
    b.empty = true; b.batch = 0; b.hs = b.he = b.ws = b.we = 0;
    if (valid) b = roi_bin(V.rois + (size_t)n * 5, V.scale, V.H, V.W, PH, PW, ph, pw);
    const float* img = V.data + (size_t)b.batch * V.H * V.W * C;
    for (int c0 = threadIdx.x * VEC; c0 < C; c0 += blockDim.x * VEC) {
        float mv[VEC];
        int mi[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) { mv[e] = b.empty ? 0.f : -FLT_MAX; mi[e] = -1; }
        for (int h = b.hs; h < b.he; ++h)
            for (int w = b.ws; w < b.we; ++w) {
                const int idx = (h * V.W + w) * C + c0;
                float x[VEC];
                if (VEC == 4) {
                    const float4 t = *reinterpret_cast<const float4*>(img + idx);
                    x[0] = t.x; x[1 % VEC] = t.y; x[2 % VEC] = t.z; x[3 % VEC] = t.w;
                } else {
                    x[0] = img[idx];
                }
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                    if (x[e] > mv[e]) { mv[e] = x[e]; mi[e] = idx + e; }
            }
        if (V.top) {
            if (VEC == 4) *reinterpret_cast<float4*>(V.top + out_base + c0) = make_float4(mv[0], mv[1 % VEC], mv[2 % VEC], mv[3 % VEC]);
            else V.top[out_base + c0] = mv[0];
        }
        if (V.argmax) {
            if (VEC == 4) *reinterpret_cast<int4*>(V.argmax + out_base + c0) = make_int4(mi[0], mi[1 % VEC], mi[2 % VEC], mi[3 % VEC]);
            else V.argmax[out_base + c0] = mi[0];
        }
        if (V.top_hi) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                __nv_bfloat16 hi, lo;
                split_bf16(mv[e], hi, lo);
                V.top_hi[out_base + c0 + e] = hi;
                if (V.top_lo) V.top_lo[out_base + c0 + e] = lo;
            }
        }
    }
}

// Backward = scatter-add of top_diff through argmax (equivalent to the reference's gather over all rois,
// roi_pooling_op_gpu.cu.cc:114-190; fp32 sum order differs -> compare at 1e-5).  bottom_diff is zeroed first.
// The reference only visits pixels inside [roi_start, roi_end] (roi_pooling_op.cc:398-409), so a malformed ROI
// (end < start, pooled as 1x1 in the forward pass) receives no gradient -- reproduced by the bounds test below.
__global__ void roi_pool_bwd_kernel(const float* __restrict__ top_diff, const int* __restrict__ argmax,
                                    const float* __restrict__ rois, long long total, int per_roi, long long img_elems,
                                    int batch_size, int W, int C, float scale, float* __restrict__ bottom_diff) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int a = argmax[i];
        if (a < 0) continue;
        const int n = (int)(i / per_roi);
        const float* roi = rois + (size_t)n * 5;
        const int b = (int)roi[0];
        if (b < 0 || b >= batch_size) continue;
        const int pix = a / C, w = pix % W, h = pix / W;
        const int rsw = (int)roundf(roi[1] * scale), rsh = (int)roundf(roi[2] * scale);
        const int rew = (int)roundf(roi[3] * scale), reh = (int)roundf(roi[4] * scale);
        if (!(w >= rsw && w <= rew && h >= rsh && h <= reh)) continue;
        atomicAdd(bottom_diff + (size_t)b * img_elems + a, top_diff[i]);
    }
}

// ---- fused multi-view kernel -------------------------------------------------------------------------------------
constexpr int kFusedThreads = 256;
constexpr int kSliceC = 64;              // channels per CTA (grid.z = C / 64): 8x the CTAs of one-per-roi
constexpr int kStageBytes = 27 * 1024;   // shared-memory window slice: 108 cells x 64 channels; 8 CTAs per SM

__device__ __forceinline__ uint4 pack8_bf16(const __nv_bfloat16* v) {
    uint4 r;
    r.x = (uint32_t)__bfloat16_as_ushort(v[0]) | ((uint32_t)__bfloat16_as_ushort(v[1]) << 16);
    r.y = (uint32_t)__bfloat16_as_ushort(v[2]) | ((uint32_t)__bfloat16_as_ushort(v[3]) << 16);
    r.z = (uint32_t)__bfloat16_as_ushort(v[4]) | ((uint32_t)__bfloat16_as_ushort(v[5]) << 16);
    r.w = (uint32_t)__bfloat16_as_ushort(v[6]) | ((uint32_t)__bfloat16_as_ushort(v[7]) << 16);
    return r;
}

// 8 consecutive channels of feature-map cell (h, w) of frame `batch` as float32: from the dense NHWC map, or rebuilt from
// the PAD operand planes (f16e5: fp16 + e5m2 residual / 4096; bf16 pair: hi + lo -- both sums are exact in float32).
__device__ __forceinline__ void load_cell8(const RoiViewDev& V, int batch, int h, int w, int C, int c0, float* x) {
    if (V.pad_hi == nullptr) {
        const float* p = V.data + (((size_t)batch * V.H + h) * V.W + w) * C + c0;
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
        return;
    }
    const size_t pix = ((size_t)batch * (V.H + 1) + h) * (V.W + 1) + w + 1;
    const uint4 hv = __ldg(reinterpret_cast<const uint4*>(V.pad_hi + pix * V.pad_c + c0));
    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
    if (V.pad_fmt == MV3D_FMT_F16E5) {
        const uint8_t* row = reinterpret_cast<const uint8_t*>(V.pad_lo) + pix * V.pad_c * 2 + f16e5_off(c0) + 64;
        const uint2 lv = __ldg(reinterpret_cast<const uint2*>(row));
        const uint32_t lw[2] = {lv.x, lv.y};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const unsigned short h16 = (unsigned short)(hw[e >> 1] >> ((e & 1) * 16));
            const uint8_t l8 = (uint8_t)(lw[e >> 2] >> ((e & 3) * 8));
            x[e] = join_f16e5(h16, l8);
        }
    } else {
        const uint4 lv = __ldg(reinterpret_cast<const uint4*>(V.pad_lo + pix * V.pad_c + c0));
        const uint32_t lw[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const uint32_t hb = (hw[e >> 1] >> ((e & 1) * 16)) << 16, lb = (lw[e >> 1] >> ((e & 1) * 16)) << 16;
            x[e] = __uint_as_float(hb) + __uint_as_float(lb);
        }
    }
}

// grid (R, n_views, ceil(C / 64)), C % 8 == 0.  Rows >= *d_num_valid: zero outputs, argmax -1, zero rectangle.
// One CTA = one roi x one view x one 64-channel slice: the window's cells of that slice are staged in shared memory
// in ONE round of independent 128-bit loads (narrower channel sub-slices when the window has more than 108 cells,
// direct reads beyond 864), then 8-channel lanes render the 49 bins from shared memory.
#ifndef MV3D_ROI_MINBLOCKS
#define MV3D_ROI_MINBLOCKS 6
#endif
__global__ void __launch_bounds__(kFusedThreads, MV3D_ROI_MINBLOCKS)
roi_pool_fused_kernel(RoiViews views, RoiProjDev proj, const float* __restrict__ rois_3d, int R,
                      const int* __restrict__ d_num_valid, int C, int PH, int PW) {
    extern __shared__ float4 stage[];
    __shared__ float roi_s[5];
    const RoiViewDev& V = views.v[blockIdx.y];
    const int n = blockIdx.x, tid = threadIdx.x;
    const int c_lo = blockIdx.z * kSliceC, c_hi = min(C, c_lo + kSliceC);   // this CTA's channels
    const bool valid = d_num_valid ? (n < *d_num_valid) : true;
    __shared__ double part_s[16];
    if (tid < 32) {   // warp 0: eight lanes do the per-corner (per-coordinate) float64 work, lane 0 combines in corner order
        float roi[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (valid && V.source == MV3D_ROI_GIVEN) {
            if (tid == 0) {
#pragma unroll
                for (int q = 0; q < 5; ++q) roi[q] = V.rois[(size_t)n * 5 + q];
            }
        } else if (valid) {
            const float* p = rois_3d + (size_t)n * 7;
            roi[0] = p[0];
            const BoxExtents e = box_extents(p[1], p[2], p[3], p[4], p[5], p[6]);
            if (V.source == MV3D_ROI_BEV) {
                if (tid < 4) part_s[tid] = (double)bev_coord(proj.bev, e, tid);
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 4; ++q) roi[1 + q] = (float)part_s[q];
            } else if (V.source == MV3D_ROI_IMG) {
                if (tid < 8) {
                    float Mloc[12];
#pragma unroll
                    for (int q = 0; q < 12; ++q) Mloc[q] = proj.dM ? __ldg(proj.dM + q) : proj.M[q];
                    img_corner_uv(Mloc, e.xp, e.xm, e.yp, e.ym, e.zp, e.zm, tid, part_s[tid], part_s[8 + tid]);
                }
                __syncwarp();
                int img[4];
                img_box_from_uv(part_s, part_s + 8, img);
#pragma unroll
                for (int q = 0; q < 4; ++q) roi[1 + q] = (float)img[q];
            } else {
                if (tid < 8) fv_corner(proj.fv, e, tid, part_s[tid], part_s[8 + tid]);
                __syncwarp();
                fv_box_from_corners(proj.fv, part_s, part_s + 8, roi + 1);
            }
        }
        if (tid == 0) {
#pragma unroll
            for (int q = 0; q < 5; ++q) roi_s[q] = roi[q];
            if (V.rois_out && blockIdx.z == 0) {
#pragma unroll
                for (int q = 0; q < 5; ++q) V.rois_out[(size_t)n * 5 + q] = roi[q];
            }
        }
    }
    __syncthreads();
    const int nbins = PH * PW;
    const size_t roi_base = (size_t)n * nbins * C;
    if (!valid) {
        const int lanes = (c_hi - c_lo) / 8;
        for (int item = tid; item < nbins * lanes; item += kFusedThreads) {
            const int bin = item / lanes, lane = item - bin * lanes;
            const size_t o = roi_base + (size_t)bin * C + c_lo + lane * 8;
            if (V.top) { *reinterpret_cast<float4*>(V.top + o) = make_float4(0, 0, 0, 0); *reinterpret_cast<float4*>(V.top + o + 4) = make_float4(0, 0, 0, 0); }
            if (V.argmax) { *reinterpret_cast<int4*>(V.argmax + o) = make_int4(-1, -1, -1, -1); *reinterpret_cast<int4*>(V.argmax + o + 4) = make_int4(-1, -1, -1, -1); }
            if (V.top_hi) *reinterpret_cast<uint4*>(V.top_hi + o) = make_uint4(0, 0, 0, 0);
            if (V.top_lo) *reinterpret_cast<uint4*>(V.top_lo + o) = make_uint4(0, 0, 0, 0);
        }
        return;
    }
    // the roi on the feature map: roi_pooling_op.cc:138-150 (round half away from zero, float32 bin sizes)
    const int H = V.H, W = V.W;
    const float scale = V.scale;
    const int batch = (int)roi_s[0];
    const int rsw = (int)roundf(roi_s[1] * scale), rsh = (int)roundf(roi_s[2] * scale);
    const int rew = (int)roundf(roi_s[3] * scale), reh = (int)roundf(roi_s[4] * scale);
    const int rw = max(rew - rsw + 1, 1), rh = max(reh - rsh + 1, 1);
    const float bsh = (float)rh / (float)PH, bsw = (float)rw / (float)PW;
    // window = union of all bins (bin bounds are monotone in ph / pw)
    const int h0 = min(max(rsh, 0), H), h1 = min(max((int)ceilf((float)PH * bsh) + rsh, 0), H);
    const int w0 = min(max(rsw, 0), W), w1 = min(max((int)ceilf((float)PW * bsw) + rsw, 0), W);
    const int wh = max(h1 - h0, 0), ww = max(w1 - w0, 0);
    const int cells = wh * ww;
    // channel sub-slice: as many of this CTA's channels (multiple of 8) as fit the stage next to `cells` window cells
    const int Cmine = c_hi - c_lo;
    int Cs = Cmine;
    if (cells > 0) {
        const int fit = (kStageBytes / 4) / cells;     // floats per cell that fit
        Cs = fit >= Cmine ? Cmine : (fit / 8) * 8;
    }
    const bool staged = Cs >= 8;
    if (!staged) Cs = Cmine;                            // very large window (> 864 cells): read global memory directly
    for (int cb = c_lo; cb < c_hi; cb += Cs) {
        const int cs = min(Cs, c_hi - cb), v4 = cs / 4, lanes = cs / 8;
        if (staged && cells > 0) {
            __syncthreads();                            // previous sub-slice fully consumed
            for (int idx = tid; idx < cells * lanes; idx += kFusedThreads) {
                const int cell = idx / lanes, v = idx - cell * lanes;
                const int h = h0 + cell / ww, w = w0 + cell % ww;
                float x[8];
                load_cell8(V, batch, h, w, C, cb + v * 8, x);
                stage[cell * v4 + 2 * v] = make_float4(x[0], x[1], x[2], x[3]);
                stage[cell * v4 + 2 * v + 1] = make_float4(x[4], x[5], x[6], x[7]);
            }
            __syncthreads();
        }
        for (int item = tid; item < nbins * lanes; item += kFusedThreads) {
            const int bin = item / lanes, lane = item - bin * lanes;
            const int ph = bin / PW, pw = bin - ph * PW;
            int hs = (int)floorf((float)ph * bsh), ws = (int)floorf((float)pw * bsw);
            int he = (int)ceilf((float)(ph + 1) * bsh), we = (int)ceilf((float)(pw + 1) * bsw);
            hs = min(max(hs + rsh, 0), H); he = min(max(he + rsh, 0), H);
            ws = min(max(ws + rsw, 0), W); we = min(max(we + rsw, 0), W);
            const bool empty = (he <= hs) || (we <= ws);
            float mv[8];
            int mi[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { mv[e] = empty ? 0.f : -FLT_MAX; mi[e] = -1; }
            for (int h = hs; h < he; ++h)
                for (int w = ws; w < we; ++w) {
                    const int idx0 = (h * W + w) * C + cb + lane * 8;
                    float x[8];
                    if (staged) {
                        const float4* sp = stage + ((h - h0) * ww + (w - w0)) * v4 + lane * 2;
                        const float4 a = sp[0], b = sp[1];
                        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
                    } else {
                        load_cell8(V, batch, h, w, C, cb + lane * 8, x);
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (x[e] > mv[e]) { mv[e] = x[e]; mi[e] = idx0 + e; }
                }
            const size_t o = roi_base + (size_t)bin * C + cb + lane * 8;
            if (V.top) {
                *reinterpret_cast<float4*>(V.top + o) = make_float4(mv[0], mv[1], mv[2], mv[3]);
                *reinterpret_cast<float4*>(V.top + o + 4) = make_float4(mv[4], mv[5], mv[6], mv[7]);
            }
            if (V.argmax) {
                *reinterpret_cast<int4*>(V.argmax + o) = make_int4(mi[0], mi[1], mi[2], mi[3]);
                *reinterpret_cast<int4*>(V.argmax + o + 4) = make_int4(mi[4], mi[5], mi[6], mi[7]);
            }
            if (V.top_hi && V.top_fmt == MV3D_FMT_F16E5) {
                uint32_t h2[4];
                unsigned short h8[4], l8[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split_f16e5_x2(mv[2 * e], mv[2 * e + 1], h2[e], h8[e], l8[e]);
                *reinterpret_cast<uint4*>(V.top_hi + o) = make_uint4(h2[0], h2[1], h2[2], h2[3]);
                const size_t k0 = (size_t)bin * C + cb + lane * 8;      // K index inside the ROI's row
                uint8_t* row = reinterpret_cast<uint8_t*>(V.top_lo) + 2 * roi_base + ((k0 >> 6) << 7) + (k0 & 63);
                *reinterpret_cast<uint2*>(row) = make_uint2(h8[0] | ((uint32_t)h8[1] << 16), h8[2] | ((uint32_t)h8[3] << 16));
                *reinterpret_cast<uint2*>(row + 64) = make_uint2(l8[0] | ((uint32_t)l8[1] << 16), l8[2] | ((uint32_t)l8[3] << 16));
            } else if (V.top_hi) {
                __nv_bfloat16 hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split_bf16(mv[e], hi[e], lo[e]);
                *reinterpret_cast<uint4*>(V.top_hi + o) = pack8_bf16(hi);
                if (V.top_lo) *reinterpret_cast<uint4*>(V.top_lo + o) = pack8_bf16(lo);
            }
        }
    }
}

static int launch_fwd(const RoiViews& views, int n_views, int R, const int* d_num_valid, int C, int PH, int PW,
                      cudaStream_t s) {
    if (R == 0) return MV3D_OK;
    dim3 grid(R * PH * PW, n_views);
    if (C % 4 == 0) {
        int threads = C / 4;
        threads = threads > 256 ? 256 : (threads < 32 ? 32 : (threads + 31) / 32 * 32);
        roi_pool_fwd_kernel<4><<<grid, threads, 0, s>>>(views, R, d_num_valid, C, PH, PW);
    } else {
        const int threads = C > 256 ? 256 : (C + 31) / 32 * 32;
        roi_pool_fwd_kernel<1><<<grid, threads, 0, s>>>(views, R, d_num_valid, C, PH, PW);
    }
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

}  // namespace mv3d

using namespace mv3d;

static int fill_views(const mv3d_roi_view* views, int n_views, bool fused, RoiViews* v) {
    for (int i = 0; i < n_views; ++i) {
        const int src = fused ? views[i].source : MV3D_ROI_GIVEN;
        MV3D_REQUIRE((views[i].d_data || views[i].d_pad_hi) && views[i].height > 0 && views[i].width > 0);
        MV3D_REQUIRE(!views[i].d_pad_hi || (views[i].d_pad_lo && views[i].pad_c % 8 == 0 &&
                                             (views[i].pad_fmt == MV3D_FMT_BF16X2 || (views[i].pad_fmt == MV3D_FMT_F16E5 && views[i].pad_c % 64 == 0))));
        MV3D_REQUIRE(src >= MV3D_ROI_GIVEN && src <= MV3D_ROI_FV);
        MV3D_REQUIRE(src != MV3D_ROI_GIVEN || views[i].d_rois);
        MV3D_REQUIRE(views[i].d_top || views[i].d_top_hi);
        v->v[i].data = views[i].d_data; v->v[i].rois = views[i].d_rois; v->v[i].H = views[i].height;
        v->v[i].W = views[i].width; v->v[i].scale = views[i].spatial_scale; v->v[i].top = views[i].d_top;
        v->v[i].argmax = views[i].d_argmax;
        v->v[i].top_hi = static_cast<__nv_bfloat16*>(views[i].d_top_hi);
        v->v[i].top_lo = static_cast<__nv_bfloat16*>(views[i].d_top_lo);
        v->v[i].source = src;
        v->v[i].rois_out = fused ? views[i].d_rois_out : nullptr;
        v->v[i].pad_hi = static_cast<const unsigned short*>(views[i].d_pad_hi);
        v->v[i].pad_lo = static_cast<const unsigned short*>(views[i].d_pad_lo);
        v->v[i].pad_fmt = views[i].pad_fmt; v->v[i].pad_c = views[i].pad_c;
        MV3D_REQUIRE(views[i].top_fmt == MV3D_FMT_BF16X2 || (views[i].top_fmt == MV3D_FMT_F16E5 && views[i].d_top_hi && views[i].d_top_lo));
        v->v[i].top_fmt = views[i].top_fmt;
    }
    return MV3D_OK;
}

static int launch_fused(const RoiViews& v, int n_views, const RoiProjDev& proj, const float* d_rois_3d, int R,
                        const int* d_num_valid, int C, int PH, int PW, cudaStream_t s) {
    if (R == 0) return MV3D_OK;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(roi_pool_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStageBytes);
        if (e != cudaSuccess) { set_last_cuda_error(e); return MV3D_ERR_LAUNCH; }
        attr_set = true;
    }
    roi_pool_fused_kernel<<<dim3(R, n_views, ceil_div(C, kSliceC)), kFusedThreads, kStageBytes, s>>>(v, proj, d_rois_3d, R, d_num_valid, C, PH, PW);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

extern "C" __attribute__((visibility("default"))) int mv3d_roi_pool_forward(
    const float* d_bottom_data, float spatial_scale, int num_rois, int height, int width, int channels,
    int pooled_height, int pooled_width, const float* d_bottom_rois, float* d_top_data, int* d_argmax_data,
    void* stream) {
    MV3D_REQUIRE(num_rois >= 0 && height > 0 && width > 0 && channels > 0 && pooled_height > 0 && pooled_width > 0);
    MV3D_REQUIRE(num_rois == 0 || (d_bottom_data && d_bottom_rois && d_top_data));
    RoiViews v = {};
    v.v[0].data = d_bottom_data; v.v[0].rois = d_bottom_rois; v.v[0].H = height; v.v[0].W = width;
    v.v[0].scale = spatial_scale; v.v[0].top = d_top_data; v.v[0].argmax = d_argmax_data;
    v.v[0].source = MV3D_ROI_GIVEN;
    if (channels % 8 == 0) {
        RoiProjDev proj = {};
        return launch_fused(v, 1, proj, nullptr, num_rois, nullptr, channels, pooled_height, pooled_width,
                            (cudaStream_t)stream);
    }
    return launch_fwd(v, 1, num_rois, nullptr, channels, pooled_height, pooled_width, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int mv3d_roi_pool_multiview(
    const mv3d_roi_view* views, int n_views, int num_rois, const int* d_num_valid, int channels, int pooled_height,
    int pooled_width, void* stream) {
    MV3D_REQUIRE(views && n_views >= 1 && n_views <= kMaxViews && num_rois >= 0 && channels > 0);
    MV3D_REQUIRE(pooled_height > 0 && pooled_width > 0);
    RoiViews v = {};
    const int rc = fill_views(views, n_views, false, &v);
    if (rc != MV3D_OK) return rc;
    for (int i = 0; i < n_views; ++i) {
        MV3D_REQUIRE(!v.v[i].pad_hi || channels % 8 == 0);                       // PAD input: staged kernel only
        MV3D_REQUIRE(v.v[i].top_fmt != MV3D_FMT_F16E5 || channels % 64 == 0);    // f16e5 operand: whole 64-channel chunks
    }
    if (channels % 8 == 0) {   // staged kernel, given rectangles
        RoiProjDev proj = {};
        return launch_fused(v, n_views, proj, nullptr, num_rois, d_num_valid, channels, pooled_height, pooled_width,
                            (cudaStream_t)stream);
    }
    return launch_fwd(v, n_views, num_rois, d_num_valid, channels, pooled_height, pooled_width, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int mv3d_roi_pool_fused(
    const mv3d_roi_view* views, int n_views, const float* d_rois_3d, const mv3d_roi_projection* proj, int num_rois,
    const int* d_num_valid, int channels, int pooled_height, int pooled_width, void* stream) {
    MV3D_REQUIRE(views && n_views >= 1 && n_views <= kMaxViews && num_rois >= 0 && channels > 0 && channels % 8 == 0);
    MV3D_REQUIRE(pooled_height > 0 && pooled_width > 0);
    RoiViews v = {};
    const int rc = fill_views(views, n_views, true, &v);
    if (rc != MV3D_OK) return rc;
    bool projected = false;
    for (int i = 0; i < n_views; ++i) {
        projected |= (v.v[i].source != MV3D_ROI_GIVEN);
        MV3D_REQUIRE(v.v[i].top_fmt != MV3D_FMT_F16E5 || channels % 64 == 0);
    }
    MV3D_REQUIRE(!projected || (d_rois_3d && proj));
    RoiProjDev pd = {};
    if (proj) {
        pd.bev.xn = proj->xn; pd.bev.yn = proj->yn; pd.bev.x_min = proj->x_min; pd.bev.y_min = proj->y_min;
        pd.bev.res = proj->res;
        pd.bev.clip_x = proj->im_w - 1.0f;   // im_shape[1] - 1 on a float32 array element (as the proposal layer)
        pd.bev.clip_y = proj->im_h - 1.0f;
        for (int q = 0; q < 12; ++q) pd.M[q] = proj->h_proj[q];
        pd.dM = proj->d_proj;
        pd.fv.H = proj->fv_h; pd.fv.W = proj->fv_w; pd.fv.theta_min = proj->fv_theta_min; pd.fv.dtheta = proj->fv_dtheta;
        pd.fv.phi_max = proj->fv_phi_max; pd.fv.dphi = proj->fv_dphi;
    }
    return launch_fused(v, n_views, pd, d_rois_3d, num_rois, d_num_valid, channels, pooled_height, pooled_width,
                        (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int mv3d_roi_pool_backward(
    const float* d_top_diff, float spatial_scale, int batch_size, int num_rois, int height, int width, int channels,
    int pooled_height, int pooled_width, const float* d_bottom_rois, float* d_bottom_diff, const int* d_argmax_data,
    void* stream) {
    MV3D_REQUIRE(batch_size > 0 && num_rois >= 0 && height > 0 && width > 0 && channels > 0 && d_bottom_diff);
    cudaStream_t s = (cudaStream_t)stream;
    const long long img = (long long)height * width * channels;
    cudaError_t e = cudaMemsetAsync(d_bottom_diff, 0, sizeof(float) * (size_t)batch_size * img, s);
    if (e != cudaSuccess) { set_last_cuda_error(e); return MV3D_ERR_LAUNCH; }
    if (num_rois == 0) return MV3D_OK;
    MV3D_REQUIRE(d_top_diff && d_bottom_rois && d_argmax_data);
    const int per_roi = pooled_height * pooled_width * channels;
    const long long total = (long long)num_rois * per_roi;
    long long g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    roi_pool_bwd_kernel<<<(int)g, 256, 0, s>>>(d_top_diff, d_argmax_data, d_bottom_rois, total, per_roi, img, batch_size,
                                              width, channels, spatial_scale, d_bottom_diff);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}
