// Greedy NMS entirely on the device.  Replaces lib/nms/cpu_nms.pyx:17-68 (rule '>=' in double) and
// lib/nms/nms_kernel.cu:34-144 (rule '>' in float; there the keep-chain is reduced on the HOST after an
// 18 MB mask D2H -- here it never leaves the GPU).
//
//  nms_mask_kernel  : upper-triangular 64x64 tiles, one 64-bit suppression word per (box, column tile).
//  nms_reduce_kernel: ONE CTA walks the keep chain 64 boxes at a time: thread 0 resolves the diagonal
//                     tile serially in registers, then all threads OR the kept rows into the shared
//                     `removed` bitmap.  Stops as soon as max_keep survivors exist (the reference
//                     computes all survivors and slices [:post_nms_topN]; the prefix is identical).
//
// IoU arithmetic is the reference's, float32 with IEEE roundings (file compiled with --fmad=false):
//   area = (x2-x1+1)*(y2-y1+1); w = max(0, min(x2)-max(x1)+1); ovr = w*h / (area_i + area_j - w*h).
#include <stdio.h>

#include "common.cuh"

namespace mv3d {

constexpr int kNmsTile = 64;

__device__ __forceinline__ float box_iou(const float4 a, const float a_area, const float4 b) {
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.f, xx2 - xx1 + 1.f), h = fmaxf(0.f, yy2 - yy1 + 1.f);
    const float inter = w * h;
    const float b_area = (b.z - b.x + 1.f) * (b.w - b.y + 1.f);
    return inter / (a_area + b_area - inter);
}

__device__ __forceinline__ float4 load_box(const float* boxes, int stride, int i) {
    if (stride == 4) return *reinterpret_cast<const float4*>(boxes + (size_t)i * 4);
    const float* p = boxes + (size_t)i * stride;
    return make_float4(p[0], p[1], p[2], p[3]);
}

__global__ void __launch_bounds__(kNmsTile)
nms_mask_kernel(const float* __restrict__ boxes, int n_max, int stride, const int* __restrict__ d_n, double thresh,
                int rule_ge, int nwords, unsigned long long* __restrict__ mask) {
    const int col_blk = blockIdx.x, row_blk = blockIdx.y;
    if (col_blk < row_blk) return;
    int n = n_max;
    if (d_n) n = min(n, *d_n);
    if (row_blk * kNmsTile >= n || col_blk * kNmsTile >= n) return;
    __shared__ float4 cbox[kNmsTile];
    const int t = threadIdx.x;
    const int cj = col_blk * kNmsTile + t;
    if (cj < n) cbox[t] = load_box(boxes, stride, cj);
    __syncthreads();
    const int i = row_blk * kNmsTile + t;
    if (i >= n) return;
    const float4 a = load_box(boxes, stride, i);
    const float a_area = (a.z - a.x + 1.f) * (a.w - a.y + 1.f);
    const int ncol = min(kNmsTile, n - col_blk * kNmsTile);
    const float thresh_f = (float)thresh;
    unsigned long long bits = 0;
    const int start = (row_blk == col_blk) ? t + 1 : 0;
    for (int j = start; j < ncol; ++j) {
        const float ovr = box_iou(a, a_area, cbox[j]);
        const bool sup = rule_ge ? ((double)ovr >= thresh) : (ovr > thresh_f);
        if (sup) bits |= 1ull << j;
    }
    mask[(size_t)i * nwords + col_blk] = bits;
}

constexpr int kReduceThreads = 1024;
constexpr int kMaxWords = 1024;  // up to 65536 boxes

__global__ void __launch_bounds__(kReduceThreads)
nms_reduce_kernel(const unsigned long long* __restrict__ mask, int n_max, const int* __restrict__ d_n, int nwords,
                  int max_keep, int* __restrict__ keep_out, int* __restrict__ num_out) {
    __shared__ unsigned long long removed[kMaxWords];
    __shared__ unsigned long long diag[kNmsTile];
    __shared__ unsigned long long keepmask_s;
    __shared__ int count_s;
    int n = n_max;
    if (d_n) n = min(n, *d_n);
    if (max_keep <= 0 || max_keep > n) max_keep = n;
    const int tid = threadIdx.x;
    for (int i = tid; i < nwords; i += blockDim.x) removed[i] = 0;
    if (tid == 0) count_s = 0;
    __syncthreads();
    const int nblk = (n + kNmsTile - 1) / kNmsTile;
    const int wlane = tid & 255, kq = tid >> 8;  // 256 word lanes x 4 row groups
    for (int b = 0; b < nblk; ++b) {
        if (tid < kNmsTile) {
            const int row = b * kNmsTile + tid;
            diag[tid] = row < n ? mask[(size_t)row * nwords + b] : ~0ull;
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long cur = removed[b], km = 0;
            int cnt = count_s;
            const int lim = min(kNmsTile, n - b * kNmsTile);
            for (int k = 0; k < lim && cnt < max_keep; ++k) {
                if (!((cur >> k) & 1ull)) {
                    km |= 1ull << k;
                    keep_out[cnt++] = b * kNmsTile + k;
                    cur |= diag[k];
                }
            }
            keepmask_s = km;
            count_s = cnt;
        }
        __syncthreads();
        if (count_s >= max_keep) break;
        const unsigned long long km = keepmask_s;
        for (int w = b + 1 + wlane; w < nwords; w += 256) {
            unsigned long long acc = 0;
#pragma unroll 4
            for (int k = kq; k < kNmsTile; k += 4)
                if ((km >> k) & 1ull) acc |= mask[(size_t)(b * kNmsTile + k) * nwords + w];
            if (acc) atomicOr(&removed[w], acc);
        }
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) *num_out = count_s;
}

static size_t nms_words(int n) { return (size_t)ceil_div(n > 0 ? n : 1, kNmsTile); }

}  // namespace mv3d

using namespace mv3d;

extern "C" __attribute__((visibility("default"))) size_t mv3d_nms_workspace_bytes(int n_boxes) {
    return align_up((size_t)(n_boxes > 0 ? n_boxes : 1) * nms_words(n_boxes) * sizeof(unsigned long long), 256);
}

extern "C" __attribute__((visibility("default"))) int mv3d_nms(const float* d_boxes, int n_boxes, int box_stride,
                                                               const int* d_n_boxes, double thresh, int rule_ge,
                                                               int max_keep, int* d_keep_out, int* d_num_out,
                                                               void* d_workspace, size_t workspace_bytes,
                                                               void* stream) {
    MV3D_REQUIRE(n_boxes >= 0 && box_stride >= 4 && d_keep_out && d_num_out);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n_boxes == 0) {  // nms_wrapper.py:16-17: empty in, empty out
        cudaError_t e = cudaMemsetAsync(d_num_out, 0, sizeof(int), s);
        if (e != cudaSuccess) { set_last_cuda_error(e); return MV3D_ERR_LAUNCH; }
        return MV3D_OK;
    }
    MV3D_REQUIRE(d_boxes != nullptr);
    const int nwords = (int)nms_words(n_boxes);
    MV3D_REQUIRE(nwords <= kMaxWords);
    if (!d_workspace || workspace_bytes < mv3d_nms_workspace_bytes(n_boxes)) return MV3D_ERR_WORKSPACE;
    unsigned long long* mask = static_cast<unsigned long long*>(d_workspace);
    dim3 grid(nwords, nwords);
    nms_mask_kernel<<<grid, kNmsTile, 0, s>>>(d_boxes, n_boxes, box_stride, d_n_boxes, thresh, rule_ge, nwords, mask);
    nms_reduce_kernel<<<1, kReduceThreads, 0, s>>>(mask, n_boxes, d_n_boxes, nwords, max_keep, d_keep_out, d_num_out);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

// Literal drop-in for the reference C ABI (lib/nms/gpu_nms.hpp:1-2): host pointers, synchronous, `>` rule.
// Errors are reported on stderr and swallowed like the reference does (nms_kernel.cu:12-19).
extern "C" __attribute__((visibility("default"))) void _nms(int* keep_out, int* num_out, const float* boxes_host,
                                                            int boxes_num, int boxes_dim, float nms_overlap_thresh,
                                                            int device_id) {
    *num_out = 0;
    if (boxes_num <= 0) return;
    int prev = 0;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device_id) != cudaSuccess) { fprintf(stderr, "_nms: bad device %d\n", device_id); return; }
    float* d_boxes = nullptr;
    int* d_keep = nullptr;
    void* d_ws = nullptr;
    const size_t ws = mv3d_nms_workspace_bytes(boxes_num);
    bool ok = cudaMalloc(&d_boxes, sizeof(float) * (size_t)boxes_num * boxes_dim) == cudaSuccess &&
              cudaMalloc(&d_keep, sizeof(int) * ((size_t)boxes_num + 1)) == cudaSuccess &&
              cudaMalloc(&d_ws, ws) == cudaSuccess;
    if (ok) {
        cudaMemcpy(d_boxes, boxes_host, sizeof(float) * (size_t)boxes_num * boxes_dim, cudaMemcpyHostToDevice);
        const int rc = mv3d_nms(d_boxes, boxes_num, boxes_dim, nullptr, (double)nms_overlap_thresh, 0, 0, d_keep + 1,
                                d_keep, d_ws, ws, nullptr);
        if (rc == MV3D_OK && cudaDeviceSynchronize() == cudaSuccess) {
            cudaMemcpy(num_out, d_keep, sizeof(int), cudaMemcpyDeviceToHost);
            cudaMemcpy(keep_out, d_keep + 1, sizeof(int) * (size_t)(*num_out), cudaMemcpyDeviceToHost);
        } else {
            fprintf(stderr, "_nms: CUDA failure: %s\n", cudaGetErrorString(cudaGetLastError()));
        }
    } else {
        fprintf(stderr, "_nms: cudaMalloc failed\n");
    }
    cudaFree(d_boxes); cudaFree(d_keep); cudaFree(d_ws);
    cudaSetDevice(prev);
}
