// Training target layers on the device.
//   anchor targets   : lib/rpn_msr/anchor_target_layer_tf.py:21-250 (+ bbox_overlaps lib/utils/bbox.pyx:15-55,
//                      bbox_transform_3d lib/fast_rcnn/bbox_transform.py:32-58)
//   proposal targets : lib/rpn_msr/proposal_target_layer_tf.py:19-94,227-298 (+ bbox_transform_cnr :61-72,
//                      lidar_3d_to_corners / lidar_cnr_to_img lib/utils/transform.py:290-315,483-500)
// The IoU matrices are float64 with the +1 pixel convention exactly like bbox.pyx (compiled with --fmad=false).
// The reference's random sub-sampling draws from numpy's global MT19937 stream (npr.choice); that step stays on
// the host in the Python mirror (it is inherently sequential and must consume the same stream to be reproducible),
// operating on the per-anchor / per-roi codes these kernels produce.
#include "common.cuh"
#include "geom.cuh"

namespace mv3d {

// bbox.pyx:35-54 for one (box, query) pair; boxes are (x1,y1,x2,y2) doubles.
__device__ __forceinline__ double iou_f64(double b0, double b1, double b2, double b3, double q0, double q1, double q2,
                                          double q3) {
    const double box_area = (q2 - q0 + 1) * (q3 - q1 + 1);
    const double iw = fmin(b2, q2) - fmax(b0, q0) + 1;
    if (iw > 0) {
        const double ih = fmin(b3, q3) - fmax(b1, q1) + 1;
        if (ih > 0) {
            const double ua = (b2 - b0 + 1) * (b3 - b1 + 1) + box_area - iw * ih;
            return iw * ih / ua;
        }
    }
    return 0.0;
}

__device__ __forceinline__ bool anchor_inside(const int* a, float im_h, float im_w) {
    // anchor_target_layer_tf.py:93-98 with _allowed_border = 0 (int64 anchors vs float32 im_info)
    return a[0] >= 0 && a[1] >= 0 && (float)a[2] < im_w && (float)a[3] < im_h;
}

// pass 1: per inside anchor max / argmax over the GT boxes (first maximum wins, numpy argmax), and the per-GT
// column maximum (atomicMax on the bit pattern of the non-negative double).
__global__ void anchor_overlap_kernel(const int* __restrict__ anchors, int N, const float* __restrict__ gt_bv, int G,
                                      float im_h, float im_w, double* __restrict__ max_ov, int* __restrict__ argmax,
                                      unsigned long long* __restrict__ gt_max_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int* a = anchors + (size_t)(i < N ? i : 0) * 4;
    const bool inside = i < N && anchor_inside(a, im_h, im_w);
    double best = -1.0;
    int arg = inside ? 0 : -1;
    for (int g = 0; g < G; ++g) {
        const float* q = gt_bv + (size_t)g * 5;
        unsigned long long bits = 0ull;   // non-negative doubles order like their bit patterns; 0 = no contribution
        if (inside) {
            const double ov = iou_f64(a[0], a[1], a[2], a[3], q[0], q[1], q[2], q[3]);
            if (ov > best) { best = ov; arg = g; }
            bits = (unsigned long long)__double_as_longlong(ov);
        }
        // one atomic per warp and GT box instead of one per anchor (34 800 x G updates of G addresses serialise in L2)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, bits, d);
            bits = o > bits ? o : bits;
        }
        if ((threadIdx.x & 31) == 0 && bits != 0ull) atomicMax(gt_max_bits + g, bits);
    }
    if (i < N) {
        max_ov[i] = best;
        argmax[i] = arg;
    }
}

// pass 2: labels before sub-sampling (anchor_target_layer_tf.py:104-143) and regression targets (:164-165).
// code[i]: bits 0-1 = label + 1 (0: don't care, 1: bg, 2: fg); bit 2 = inside the image; bit 3 = max_overlap <
// RPN_NEGATIVE_OVERLAP (the set the reference relabels to 0 at :176).
__global__ void anchor_label_kernel(const int* __restrict__ anchors, const double* __restrict__ anchors3d, int N,
                                    const float* __restrict__ gt_bv, const float* __restrict__ gt_3d, int G,
                                    float im_h, float im_w, double pos_thr, double neg_thr, int clobber,
                                    const double* __restrict__ max_ov, const int* __restrict__ argmax,
                                    const unsigned long long* __restrict__ gt_max_bits,
                                    signed char* __restrict__ code, float* __restrict__ targets) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float* t = targets + (size_t)i * 6;
    const int arg = argmax[i];
    if (arg < 0) {
        code[i] = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) t[k] = 0.f;
        return;
    }
    const int* a = anchors + (size_t)i * 4;
    const double mo = max_ov[i];
    int label = -1;
    if (!clobber && 0 < mo && mo < neg_thr) label = 0;               // :126-130
    bool is_gt_best = false;                                           // :123,133
    for (int g = 0; g < G; ++g) {
        const float* q = gt_bv + (size_t)g * 5;
        const double ov = iou_f64(a[0], a[1], a[2], a[3], q[0], q[1], q[2], q[3]);
        if (ov == __longlong_as_double((long long)gt_max_bits[g])) is_gt_best = true;
    }
    if (is_gt_best) label = 1;
    if (mo >= pos_thr) label = 1;                                      // :139
    if (clobber && mo < neg_thr) label = 0;                            // :141-143
    code[i] = (signed char)((label + 1) | 4 | (mo < neg_thr ? 8 : 0));
    // bbox_transform_3d (bbox_transform.py:32-58): ex = anchor (float64), gt = float32 row
    const double* ex = anchors3d + (size_t)i * 6;
    const float* gt = gt_3d + (size_t)arg * 7;
    t[0] = (float)(((double)gt[0] - ex[0]) / ex[4]);
    t[1] = (float)(((double)gt[1] - ex[1]) / ex[3]);
    t[2] = (float)(((double)gt[2] - ex[2]) / ex[5]);
    t[3] = (float)log((double)gt[3] / ex[3]);
    t[4] = (float)log((double)gt[4] / ex[4]);
    t[5] = (float)log((double)gt[5] / ex[5]);
}

// rois (R,5) [batch,x1,y1,x2,y2] followed by the G GT boxes -> max overlap / assignment per candidate
// (proposal_target_layer_tf.py:38-44,232-236).
__global__ void roi_overlap_kernel(const float* __restrict__ rois, int R, const float* __restrict__ gt_bv, int G,
                                   double* __restrict__ max_ov, int* __restrict__ argmax) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R + G) return;
    const float* b = i < R ? rois + (size_t)i * 5 + 1 : gt_bv + (size_t)(i - R) * 5;
    double best = -1.0;
    int arg = 0;
    for (int g = 0; g < G; ++g) {
        const float* q = gt_bv + (size_t)g * 5;
        const double ov = iou_f64(b[0], b[1], b[2], b[3], q[0], q[1], q[2], q[3]);
        if (ov > best) { best = ov; arg = g; }
    }
    max_ov[i] = best;
    argmax[i] = arg;
}

// Sampled rois -> the five outputs of proposal_target_layer_3d.  keep (K) indexes the candidate list (rois ++ GT);
// the first n_fg entries are foreground.
__global__ void proposal_target_kernel(const float* __restrict__ rois_bv, const float* __restrict__ rois_3d, int R,
                                       const float* __restrict__ gt_bv, const float* __restrict__ gt_3d,
                                       const float* __restrict__ gt_cnr, int G, const int* __restrict__ keep, int K,
                                       int n_fg, const int* __restrict__ assign, const float* __restrict__ M,
                                       int num_classes, float batch_index, float* __restrict__ o_bv,
                                       float* __restrict__ o_img, int* __restrict__ o_lab, float* __restrict__ o_tgt,
                                       float* __restrict__ o_3d) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= K) return;
    const int i = keep[r];
    float bv[5], p[7];
    if (i < R) {
#pragma unroll
        for (int k = 0; k < 5; ++k) bv[k] = rois_bv[(size_t)i * 5 + k];
#pragma unroll
        for (int k = 0; k < 7; ++k) p[k] = rois_3d[(size_t)i * 7 + k];
    } else {  // appended GT rows: (0, box[:-1])
        bv[0] = 0.f; p[0] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) bv[k + 1] = gt_bv[(size_t)(i - R) * 5 + k];
#pragma unroll
        for (int k = 0; k < 6; ++k) p[k + 1] = gt_3d[(size_t)(i - R) * 7 + k];
    }
    bv[0] = batch_index; p[0] = batch_index;  // frame index for the multi-frame ROI pool (0 in the reference)
    const int g = assign[i];
    int lab = (int)gt_bv[(size_t)g * 5 + 4];
    if (r >= n_fg) lab = 0;                                            // :276
#pragma unroll
    for (int k = 0; k < 5; ++k) o_bv[(size_t)r * 5 + k] = bv[k];
#pragma unroll
    for (int k = 0; k < 7; ++k) o_3d[(size_t)r * 7 + k] = p[k];
    o_lab[r] = lab;
    // corners (transform.py:305-313), float32
    const float hl = __fdiv_rn(p[4], 2.f), hw = __fdiv_rn(p[5], 2.f), hh = __fdiv_rn(p[6], 2.f);
    const float xp = __fadd_rn(hl, p[1]), xm = __fadd_rn(-hl, p[1]);
    const float yp = __fadd_rn(hw, p[2]), ym = __fadd_rn(-hw, p[2]);
    const float zp = __fadd_rn(hh, p[3]), zm = __fadd_rn(-hh, p[3]);
    float cnr[24];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const bool sx = (c == 0 || c == 1 || c == 4 || c == 5);
        const bool sy = (c == 0 || c == 3 || c == 4 || c == 7);
        cnr[c] = sx ? xp : xm;
        cnr[8 + c] = sy ? yp : ym;
        cnr[16 + c] = (c >= 4) ? zp : zm;
    }
    int img[4];
    corners_to_img_box(M, xp, xm, yp, ym, zp, zm, img);
    o_img[(size_t)r * 5] = bv[0];
#pragma unroll
    for (int k = 0; k < 4; ++k) o_img[(size_t)r * 5 + 1 + k] = (float)img[k];
    // bbox_transform_cnr (bbox_transform.py:61-72) placed at [24*cls, 24*cls+24) (proposal_target_layer_tf.py:172-194)
    float* t = o_tgt + (size_t)r * 24 * num_classes;
    for (int k = 0; k < 24 * num_classes; ++k) t[k] = 0.f;
    if (lab > 0 && lab < num_classes) {
        const float* gc = gt_cnr + (size_t)g * 25;
        const float dx = __fsub_rn(gc[0], gc[6]), dy = __fsub_rn(gc[8], gc[14]), dz = __fsub_rn(gc[16], gc[22]);
        // np.linalg.norm on a float32 (n,3) array with axis=1: sqrt(sum of squares) in float32, added in order
        const float diag = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
#pragma unroll
        for (int k = 0; k < 24; ++k) t[24 * lab + k] = __fdiv_rn(__fsub_rn(gc[k], cnr[k]), diag);
    }
}

}  // namespace mv3d

using namespace mv3d;
#define MV3D_API extern "C" __attribute__((visibility("default")))

MV3D_API int mv3d_anchor_targets(const int* d_anchors, const double* d_anchors3d, int N, const float* d_gt_bv,
                                 const float* d_gt_3d, int G, float im_h, float im_w, double pos_thr, double neg_thr,
                                 int clobber, double* d_max_ov, int* d_argmax, unsigned long long* d_gt_max_ws,
                                 signed char* d_code, float* d_targets, void* stream) {
    MV3D_REQUIRE(d_anchors && d_anchors3d && d_gt_bv && d_gt_3d && d_max_ov && d_argmax && d_gt_max_ws && d_code && d_targets);
    MV3D_REQUIRE(N > 0 && G > 0);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_gt_max_ws, 0, sizeof(unsigned long long) * G, s);
    if (e != cudaSuccess) { set_last_cuda_error(e); return MV3D_ERR_LAUNCH; }
    anchor_overlap_kernel<<<ceil_div(N, 128), 128, 0, s>>>(d_anchors, N, d_gt_bv, G, im_h, im_w, d_max_ov, d_argmax,
                                                           d_gt_max_ws);
    anchor_label_kernel<<<ceil_div(N, 128), 128, 0, s>>>(d_anchors, d_anchors3d, N, d_gt_bv, d_gt_3d, G, im_h, im_w,
                                                         pos_thr, neg_thr, clobber, d_max_ov, d_argmax, d_gt_max_ws,
                                                         d_code, d_targets);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

MV3D_API int mv3d_roi_overlaps(const float* d_rois_bv, int R, const float* d_gt_bv, int G, double* d_max_ov,
                               int* d_argmax, void* stream) {
    MV3D_REQUIRE(d_gt_bv && d_max_ov && d_argmax && R >= 0 && G > 0 && (R == 0 || d_rois_bv));
    roi_overlap_kernel<<<ceil_div(R + G, 128), 128, 0, (cudaStream_t)stream>>>(d_rois_bv, R, d_gt_bv, G, d_max_ov,
                                                                                d_argmax);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

MV3D_API int mv3d_proposal_targets(const float* d_rois_bv, const float* d_rois_3d, int R, const float* d_gt_bv,
                                   const float* d_gt_3d, const float* d_gt_corners, int G, const int* d_keep, int K,
                                   int n_fg, const int* d_assign, const float* d_proj, int num_classes,
                                   float batch_index, float* d_out_bv, float* d_out_img, int* d_out_labels,
                                   float* d_out_targets, float* d_out_3d, void* stream) {
    MV3D_REQUIRE(d_gt_bv && d_gt_3d && d_gt_corners && d_keep && d_assign && d_proj && d_out_bv && d_out_img &&
                 d_out_labels && d_out_targets && d_out_3d);
    MV3D_REQUIRE(R >= 0 && G > 0 && K > 0 && num_classes > 0 && (R == 0 || (d_rois_bv && d_rois_3d)));
    proposal_target_kernel<<<ceil_div(K, 64), 64, 0, (cudaStream_t)stream>>>(
        d_rois_bv, d_rois_3d, R, d_gt_bv, d_gt_3d, d_gt_corners, G, d_keep, K, n_fg, d_assign, d_proj, num_classes,
        batch_index, d_out_bv, d_out_img, d_out_labels, d_out_targets, d_out_3d);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}
