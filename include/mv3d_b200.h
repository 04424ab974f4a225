/* mv3d_b200 -- C ABI of the B200-native MV3D per-frame hot path (libmv3d_b200.so).
 *
 * Conventions (every entry point):
 *   - plain pointers and sizes only; pointers named d_* are DEVICE pointers on the current CUDA
 *     device, h_* are HOST pointers; `stream` is a cudaStream_t passed as void* (NULL = default stream)
 *   - returns MV3D_OK (0) or a negative mv3d_status; mv3d_status_string() explains it
 *   - device-pointer entry points never allocate, never synchronise and never touch host memory
 *     except their scalar arguments: scratch comes from the caller (mv3d_*_workspace_bytes)
 *   - the *_host entry points mirror reference ABIs that take HOST buffers (they allocate, copy and
 *     synchronise, exactly like the reference functions they replace)
 *
 * Each declaration cites the reference interface (leeyevi/MV3D_TF, paths relative to its root) it
 * replaces.  INTEGRATION.md shows the reference-side binding for each.
 */
#ifndef MV3D_B200_H_
#define MV3D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    MV3D_OK = 0,
    MV3D_ERR_ARG = -1,       /* invalid argument (null pointer, bad shape, unsupported size) */
    MV3D_ERR_WORKSPACE = -2, /* workspace too small */
    MV3D_ERR_LAUNCH = -3,    /* CUDA launch / runtime failure, see mv3d_last_cuda_error() */
    MV3D_ERR_DRIVER = -4     /* driver entry point (cuTensorMapEncodeTiled) unavailable */
} mv3d_status;

int mv3d_version(void);
const char* mv3d_status_string(int status);
/* cudaError_t of the most recent MV3D_ERR_LAUNCH on this thread (0 if none) and its text. */
int mv3d_last_cuda_error(void);
const char* mv3d_last_cuda_error_string(void);

/* ---------------------------------------------------------------------------------------------
 * (i) LiDAR -> bird's-eye-view raster.   Replaces point_cloud_2_top, tools/read_lidar.py:10-115
 *     (= lib/utils/read_lidar.py:10-115).  Same result bit for bit: top (H,W,C) float32, HWC.
 *
 *   d_points   (n_points, point_stride) float32 rows [x, y, z, reflectance, ...]
 *   H, W, C    = y_max+1, x_max+1, z_max+1 of read_lidar.py:49-53;  nslices = len(np.arange(h0,h1,zres))
 *   h_slice_lo / h_slice_hi  HOST arrays of nslices doubles: slice i keeps lo[i] <= z < hi[i]
 *                            (read_lidar.py:80-83; float64 compare)
 *   res, fwd0, fwd1, side0, side1, height0 : the scalars of the reference call, as float32
 *   xoff = int(floor(side0/res)), yoff = int(floor(fwd1/res))   (read_lidar.py:102-103)
 *   mv3d_bev_raster_pad writes the same raster directly in the conv trunk's input layout instead:
 *   a zero-haloed bf16 hi/lo pair of shape (H+1, W+1, c_pad) (see mv3d_pad_nhwc; d_pad_lo may be NULL),
 *   saving the float32 map's write and re-read when the BEV only feeds the network.
 * ------------------------------------------------------------------------------------------- */
size_t mv3d_bev_raster_workspace_bytes(int n_points, int H, int W, int nslices);
int mv3d_bev_raster(const float* d_points, int n_points, int point_stride, float* d_top, int H, int W, int C,
                    int nslices, const double* h_slice_lo, const double* h_slice_hi, float res, float fwd0,
                    float fwd1, float side0, float side1, float height0, int xoff, int yoff, void* d_workspace,
                    size_t workspace_bytes, void* stream);
int mv3d_bev_raster_pad(const float* d_points, int n_points, int point_stride, void* d_pad_hi, void* d_pad_lo,
                        int c_pad, int H, int W, int C, int nslices, const double* h_slice_lo,
                        const double* h_slice_hi, float res, float fwd0, float fwd1, float side0, float side1,
                        float height0, int xoff, int yoff, void* d_workspace, size_t workspace_bytes, void* stream);
/* mv3d_bev_raster_pad with the PAD planes rendered in `fmt` (MV3D_FMT_BF16X2, or MV3D_FMT_F16E5 with c_pad % 64 == 0). */
int mv3d_bev_raster_pad_fmt(const float* d_points, int n_points, int point_stride, void* d_pad_hi, void* d_pad_lo,
                            int c_pad, int H, int W, int C, int nslices, const double* h_lo, const double* h_hi,
                            float res, float fwd0, float fwd1, float side0, float side1, float height0, int xoff,
                            int yoff, void* d_workspace, size_t workspace_bytes, int fmt, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (iii) NMS.
 *   _nms: literal drop-in for the reference's C ABI, lib/nms/gpu_nms.hpp:1-2 / nms_kernel.cu:91-144
 *         (HOST pointers, boxes (boxes_num, boxes_dim>=4) float32 sorted by score descending, keeps
 *         `ovr > thresh` in float like nms_kernel.cu:71, synchronous).
 *   mv3d_nms: device-pointer form.  rule_ge=1 reproduces lib/nms/cpu_nms.pyx:65 (`(double)ovr >= thresh`),
 *         rule_ge=0 reproduces nms_kernel.cu:71.  d_boxes must already be in score-descending order
 *         (gpu_nms.pyx:23-25 sorts on the host before calling _nms).  Stops after max_keep survivors
 *         (<=0: no limit), which equals the reference's `keep[:post_nms_topN]` (proposal_layer_tf.py:173).
 *         d_n_boxes (optional) overrides n_boxes with a count that lives on the device.
 *         0 < max_keep <= 2048 (the proposal layer: 300 / 2000) runs as ONE 8-CTA cluster that tests candidates against
 *         kept boxes only (nms_lazy_kernel: same predicate, same order, identical survivor list; the workspace is not
 *         touched); otherwise the all-pairs mask + keep chain.  MV3D_NMS_LAZY=0 forces the latter.
 * ------------------------------------------------------------------------------------------- */
void _nms(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim,
          float nms_overlap_thresh, int device_id);
size_t mv3d_nms_workspace_bytes(int n_boxes);
int mv3d_nms(const float* d_boxes, int n_boxes, int box_stride, const int* d_n_boxes, double thresh, int rule_ge,
             int max_keep, int* d_keep_out, int* d_num_out, void* d_workspace, size_t workspace_bytes,
             void* stream);

/* ---------------------------------------------------------------------------------------------
 * (iii) proposal_layer_3d.   Replaces lib/rpn_msr/proposal_layer_tf.py:25-202 with the helpers it
 *     calls: generate_anchors_bv (generate_anchors.py:37-51), bv_anchor_to_lidar / lidar_3d_to_bv /
 *     lidar_3d_to_corners / lidar_cnr_to_img (lib/utils/transform.py:89-142,290-315,483-500),
 *     bbox_transform_inv_3d / clip_boxes (lib/fast_rcnn/bbox_transform.py:108-155,178-191),
 *     _filter_boxes / _filter_img_boxes (:336-352), score sort (:161-167) and nms (:172).
 *
 *   d_prob (Hf,Wf,2A) float32 = rpn_cls_prob_reshape[0], d_deltas (Hf,Wf,6A) = rpn_bbox_pred[0]
 *   d_anchors3d (Hf*Wf*A, 6) float32 : bv_anchor_to_lidar(all anchors) cast to float32 (host-made once)
 *   h_proj 12 floats: row-major 3x4 (P2 . R0) . Tr  in float32 (transform.py:383-384)
 *   geometry: xn, yn, x_min, y_min, res of transform.py:3-20 (doubles);  im_h, im_w, im_scale = im_info
 *   img_h, img_w: the image size of _filter_img_boxes (the reference hard-codes 375, 1242)
 *   Outputs (capacity post_nms_top_n rows): d_blob_bv (R,5), d_blob_img (R,5), d_blob_3d (R,7) float32
 *   with leading batch index `batch_index`; d_scores (R) ; d_anchor_index (R) int32 ; d_num_out (1) int32.
 *   Rows >= *d_num_out are zero-filled.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int Hf, Wf, A;
    double xn, yn, x_min, y_min, res;
    float im_h, im_w, im_scale;
    float img_h, img_w;
    float min_size; /* cfg RPN_MIN_SIZE */
    int pre_nms_top_n, post_nms_top_n;
    double nms_thresh;
    int nms_rule_ge;
    float batch_index;
    const float* d_proj; /* optional DEVICE copy of the 12 projection floats; when set it overrides h_proj (which may
                            then be NULL) so that a captured CUDA graph can be replayed with a per-frame calib */
    int ld_prob, ld_deltas; /* floats between consecutive feature-map cells of d_prob / d_deltas; 0 = dense (2A / 6A).
                               Lets both be column ranges of one fused (Hf*Wf, 8A) head output */
} mv3d_proposal_params;

size_t mv3d_proposal_workspace_bytes(const mv3d_proposal_params* p);
int mv3d_proposal_layer_3d(const float* d_prob, const float* d_deltas, const float* d_anchors3d,
                           const float* h_proj, const mv3d_proposal_params* p, float* d_blob_bv,
                           float* d_blob_img, float* d_blob_3d, float* d_scores, int* d_anchor_index,
                           int* d_num_out, void* d_workspace, size_t workspace_bytes, void* stream);
/* Stage outputs of the decode kernel alone, for stage-wise parity tests: per anchor score, p3d (6),
 * clipped bv box (4), image box (4, int32), keep flag (uint8). */
int mv3d_proposal_decode(const float* d_prob, const float* d_deltas, const float* d_anchors3d, const float* h_proj,
                         const mv3d_proposal_params* p, float* d_score, float* d_p3d, float* d_pbv, int* d_pimg,
                         unsigned char* d_keep, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (iv) ROI pooling.  Replaces the RoiPool / RoiPoolGrad TF ops, lib/roi_pooling_layer/
 *     roi_pooling_op.cc:30-49 and their launchers ROIPoolForwardLaucher / ROIPoolBackwardLaucher
 *     (roi_pooling_op_gpu.h:18-27, roi_pooling_op_gpu.cu.cc:87-110,193-215): device pointers,
 *     NHWC float32 data, rois (R,5) [batch,x1,y1,x2,y2], top/argmax (R,PH,PW,C).
 *     The multi-view form pools up to 3 feature maps in ONE launch (north-star (iv)).
 * ------------------------------------------------------------------------------------------- */
int mv3d_roi_pool_forward(const float* d_bottom_data, float spatial_scale, int num_rois, int height, int width,
                          int channels, int pooled_height, int pooled_width, const float* d_bottom_rois,
                          float* d_top_data, int* d_argmax_data, void* stream);
int mv3d_roi_pool_backward(const float* d_top_diff, float spatial_scale, int batch_size, int num_rois, int height,
                           int width, int channels, int pooled_height, int pooled_width,
                           const float* d_bottom_rois, float* d_bottom_diff, const int* d_argmax_data,
                           void* stream);

/* where a view's rectangles come from in mv3d_roi_pool_fused */
#define MV3D_ROI_GIVEN 0 /* d_rois (R,5), as the reference op receives them */
#define MV3D_ROI_BEV 1   /* projected in the kernel from d_rois_3d: lidar_3d_to_bv + clip_boxes (transform.py:113-142) */
#define MV3D_ROI_IMG 2   /* 8 corners -> image box (transform.py:290-315,483-500), int32-truncated */
#define MV3D_ROI_FV 3    /* front-view box (this repo's specification; the reference has no FV branch) */
typedef struct {
    const float* d_data;  /* (B,H,W,C) float32 NHWC */
    const float* d_rois;  /* (R,5); may be NULL when source != MV3D_ROI_GIVEN */
    int height, width;
    float spatial_scale;
    float* d_top;         /* (R,PH,PW,C) float32, may be NULL */
    int* d_argmax;        /* (R,PH,PW,C) int32, may be NULL */
    void* d_top_hi;       /* optional bf16 (R, PH*PW*C) hi/lo pair feeding fc6, may be NULL */
    void* d_top_lo;
    int source;           /* MV3D_ROI_* (mv3d_roi_pool_fused only; mv3d_roi_pool_multiview always reads d_rois) */
    float* d_rois_out;    /* optional (R,5) [batch,x1,y1,x2,y2]: the rectangle the kernel pooled (fused form) */
    /* fused form, optional: read the feature map from the PAD operand planes the producing conv already writes
     * ((B, H+1, W+1, pad_c), pad_fmt = MV3D_FMT_*) instead of a dense float32 copy (d_data may then be NULL); the values
     * pooled are the operand renderings (what Network.run returns for that layer when fetched) */
    const void* d_pad_hi; const void* d_pad_lo; int pad_fmt, pad_c;
    /* rendering of d_top_hi / d_top_lo: MV3D_FMT_BF16X2 (bf16 hi/lo, the default 0) or MV3D_FMT_F16E5 -- d_top_hi = fp16
     * (R, PH*PW*C), d_top_lo = the byte plane of the same rows (K index = bin*C + c, 64-element chunks [e5m2(h) | e5m2
     * residual]): fc6's operand for the 2-pass fc GEMM (mv3d_gemm_desc.passes = 2).  channels % 64 == 0. */
    int top_fmt;
} mv3d_roi_view;
/* projection constants of the fused form: BEV grid (transform.py:3-20) + clip bounds (im_info), the float32 3x4 image
 * projection (P2.R0).Tr by value or as a device pointer (graph replay), the FV map geometry (radians) */
typedef struct {
    double xn, yn, x_min, y_min, res;
    float im_h, im_w;
    float h_proj[12];
    const float* d_proj;  /* overrides h_proj when not NULL */
    int fv_h, fv_w;
    double fv_theta_min, fv_dtheta, fv_phi_max, fv_dphi;
} mv3d_roi_projection;
int mv3d_roi_pool_multiview(const mv3d_roi_view* views, int n_views, int num_rois, const int* d_num_valid,
                            int channels, int pooled_height, int pooled_width, void* stream);
/* north-star (iv): ONE launch for all views that also PROJECTS each 3-D proposal (d_rois_3d (R,7) [batch,x,y,z,l,w,h])
 * into the views whose `source` says so, stages every roi window in shared memory once and emits fc6's bf16 hi/lo
 * operand with 16-byte stores.  Replaces the two RoiPool launches + proposal_transform of the reference
 * (lib/networks/MV3D_test.py:87-113, network.py:292-315).  channels % 8 == 0. */
int mv3d_roi_pool_fused(const mv3d_roi_view* views, int n_views, const float* d_rois_3d,
                        const mv3d_roi_projection* proj, int num_rois, const int* d_num_valid, int channels,
                        int pooled_height, int pooled_width, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (ii) conv / fc as one tcgen05 implicit GEMM.  Replaces Network.conv (+bias+ReLU) and Network.fc,
 *     lib/networks/network.py:108-132,369-397 (tf.nn.conv2d SAME stride 1 / xw_plus_b).
 *
 *   Activations live in a zero-haloed, channel-padded bf16 layout "PAD": (B, Hp=H+1, Wp=W+1, Cpad),
 *   pixel (h,w) at [h][w+1], column 0 and row H are zero (one halo column / row is shared by
 *   neighbours), so a 3x3 SAME conv is 9 row-shifted GEMMs over the flattened pixel index.
 *   Precision: a value x is the pair (hi=bf16(x), lo=bf16(x-hi)).  passes=3 accumulates
 *   hi*hi + lo*hi + hi*lo in fp32 (|err| ~ 2^-16, the parity mode); passes=1 uses hi only.
 *   D[m, n] = sum_{t<taps, c<Cin} A[m + shift_t, c] * W[n, t*Cin + c]  (+ bias[n], ReLU)
 * ------------------------------------------------------------------------------------------- */
#define MV3D_FMT_BF16X2 0 /* x = hi + lo, two bf16 planes */
#define MV3D_FMT_F16E5 1  /* x ~= h + l/4096: fp16 plane h; byte plane, per 64-channel chunk: 64 x e5m2(h), 64 x e5m2(l)
                             (weights: fp16 plane = fp16(4096 w); byte plane = 64 x e5m2(residual), 64 x e5m2(w)) */
typedef struct {
    int M;    /* rows of A and D: B*Hp*Wp pixels of the PAD layout, or plain rows for fc (taps=1) */
    int N;    /* output channels */
    int Cin;  /* channels per tap in A (padded to a multiple of 16) */
    int taps; /* 9 (3x3) or 1 (1x1 / fc) */
    int Hp, Wp; /* PAD geometry; 0,0 => A is a plain (M,Cin) matrix, no halo handling */
    int passes; /* 1 or 3 */
    const void* d_a_hi; const void* d_a_lo; /* bf16 (M, Cin) */
    const void* d_w_hi; const void* d_w_lo; /* bf16 (N, taps*Cin), see mv3d_pack_conv_weights */
    const float* d_bias;                    /* (N) or NULL */
    int relu;
    void* d_out_hi; void* d_out_lo; int ld_out;  /* bf16 (M, ld_out) PAD layout (halo pixels written as 0); NULL to skip */
    float* d_out_f32; int ld_f32; int f32_dense; /* fp32 output; f32_dense=1 drops halo pixels: (B,H,W,ld_f32) */
    int split_k; /* >1: K range split over blockIdx.z, partial sums atomically added into d_out_f32
                    (which the caller zeroed); bias/relu/bf16 outputs are then ignored */
    /* backward-data epilogue (training): out = (acc + addend) * (mask > 0 ? mask_scale : 0).
     * d_mask_hi: bf16, same row indexing as d_out_hi with pitch ld_mask -- the forward activation whose ReLU
     * (and dropout) mask gates this gradient; d_addend_f32: dense float32 (B,H,W,ld_addend) (f32_dense row
     * indexing) or plain rows when Hp == 0 -- a second gradient path summed in before masking.  NULL = unused. */
    const void* d_mask_hi; int ld_mask; float mask_scale;
    const float* d_addend_f32; int ld_addend;
    /* operand rendering of d_out_hi / d_out_lo: MV3D_FMT_BF16X2 (bf16 hi/lo planes, the default 0) or MV3D_FMT_F16E5
     * (fp16 plane + e5m2 byte plane, below).  passes=2 declares that A and W are MV3D_FMT_F16E5 operands
     * (3x3 convs with Cin % 64 == 0 only): one fp16 pass + one e5m2 pass at twice the rate carrying both first-order
     * correction terms -- 2/3 of the tensor-pipe time of passes=3, error ~2^-14.5 per product (3-pass: 2^-17). */
    int out_fmt;
    /* > 0: columns [0, softmax_cols) of the float32 output are adjacent (bg, fg) score pairs; the epilogue replaces
     * each pair by its softmax (the reference's reshape -> softmax -> reshape over the last dim of 2,
     * lib/networks/MV3D_test.py:76-81, network.py:399-405) -- used to fold rpn_cls_score | rpn_bbox_pred into ONE
     * N = 32 GEMM whose output is (prob x 8 | deltas x 24).  Requires d_out_f32, relu = 0, split_k <= 1, even. */
    int softmax_cols;
    /* 1: Network.max_pool(2,2,2,2,'VALID') (network.py:181-188) of this 3x3 conv's output is taken in the epilogue
     * (CTA pair = 2 image rows x 128 columns, row maxima exchanged through distributed shared memory); d_out_hi / d_out_lo
     * are then the POOLED PAD planes (B, H/2 + 1, W/2 + 1, ld_out), halos zeroed by the kernel; the un-pooled activation
     * is never written.  N = 64 or 128, Cin % 64 == 0, passes 2 or 3, no float32 / gate / addend output. */
    int pool;
    /* > 0 with Cin == 64: only the first cin_valid input channels of the (zero-padded) 64-channel chunk are non-zero
     * (the BEV map: 36), so the CTA-pair 3x3 kernel skips the 16-channel k-steps that would multiply zeros.  0 = all. */
    int cin_valid;
} mv3d_gemm_desc;
int mv3d_conv_gemm(const mv3d_gemm_desc* desc, void* stream);
/* A/B switch for the CTA-pair (tcgen05 cta_group::2, 256 x N tiles) form of the tap-reuse 3x3 conv kernel: on (default,
 * also MV3D_PAIR=1) or off (single-CTA 128 x N tiles).  Same results either way; returns the previous setting. */
int mv3d_gemm_set_pair_mode(int on);
/* Measurement only: a device buffer of 8 int64 that CTA pair 0 of every following 3x3 pair-kernel launch fills with
 * clock64() at its phase boundaries (start, set-up done, first operands landed, last MMA issued, last accumulator
 * complete, epilogue done, both CTAs done, exit); NULL switches it off.  tools/gemm_phases.py prints the breakdown. */
int mv3d_gemm_set_stamps(void* d_stamps8);

/* MV3D_FMT_F16E5 renderings of the layout helpers (fmt = MV3D_FMT_BF16X2 forwards to the plain functions; same
 * arguments otherwise).  d_hi is the 16-bit plane, d_lo the byte plane of equal pitch; c_pad / cin_pad % 64 == 0. */
int mv3d_pack_weights_fmt(const float* d_w, int taps, int cin, int cout, int cin_pad, void* d_hi, void* d_lo, int fmt,
                          void* stream);
int mv3d_pad_nhwc_fmt(const float* d_in, int B, int H, int W, int C, int c_pad, void* d_hi, void* d_lo, int fmt,
                      void* stream);
int mv3d_unpad_nhwc_fmt(const void* d_hi, const void* d_lo, int B, int H, int W, int C, int c_pad, float* d_out, int fmt,
                        void* stream);
int mv3d_maxpool2x2_pad_fmt(const void* d_in_hi, const void* d_in_lo, int B, int H, int W, int c_pad, void* d_out_hi,
                            void* d_out_lo, int fmt, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backward-filter GEMM (training):  dW[t, c, n] (+)= sum_p X[p + shift_t, c] * G[p, n].
 *   Replaces the filter-gradient TensorFlow derives for Network.conv / Network.fc when the reference calls
 *   tf.train.AdamOptimizer(lr).minimize(loss)  (lib/fast_rcnn/train_mv.py:144-146; ops at network.py:114,395).
 *   X = the layer's forward input, G = dLoss/d(pre-bias output) with the ReLU mask applied, both bf16 hi/lo in
 *   the PAD layout (taps == 9, Wp = PAD row width) or plain rows (taps == 1).  d_dw is float32
 *   (taps, cin, cout) = the reference's HWIO / (in, out) variable layout.  accumulate=1: dW += (row splits use
 *   red.global.add; caller zeroes once per step); accumulate=0: dW = (single split, plain stores).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int P;        /* rows of X and G: B*Hp*Wp PAD pixels, or plain rows (ROIs) for fc */
    int Cx, Cg;   /* row pitch in elements (padded channels) of X (16 or multiple of 64) and G (multiple of 64) */
    int cin, cout;
    int taps;     /* 9 or 1 */
    int Wp;
    int passes;   /* 1 or 3 (bf16 hi/lo split, see mv3d_gemm_desc) */
    const void* d_x_hi; const void* d_x_lo;
    const void* d_g_hi; const void* d_g_lo;
    float* d_dw; int ld_dw; /* ld_dw = 0 -> cout */
    int accumulate;
    int split_rows;  /* 0 = automatic */
    int tap_window;  /* 1: the three kw taps of a kernel row share one displaced X window box; 0: one box per tap */
} mv3d_wgrad_desc;
int mv3d_conv_wgrad(const mv3d_wgrad_desc* desc, void* stream);

/* HWIO float32 weights (kh,kw,Cin,Cout) (network.py:119) -> bf16 hi/lo (Cout, kh*kw*cin_pad), K-major.
 * (The reference's NHWC->NCHW flatten before fc6, network.py:381, is folded into the weight by permuting
 * its rows on the caller's side before packing.) */
int mv3d_pack_weights(const float* d_w_hwio, int taps, int cin, int cout, int cin_pad, void* d_w_hi, void* d_w_lo,
                      void* stream);
/* (B,H,W,C) float32 NHWC -> PAD bf16 hi/lo (B,H+1,W+1,c_pad), halos and channel padding zeroed. */
int mv3d_pad_nhwc(const float* d_in, int B, int H, int W, int C, int c_pad, void* d_hi, void* d_lo, void* stream);
/* First-layer im2col for tiny channel counts (C = 3: RGB, front view): dense float32 (B,H,W,C) -> PAD rows
 * (B,H+1,W+1,k_pad) with K index tap*C + c (tap = kh*3+kw, SAME zero padding; k >= 9*C and halo rows zero), so that
 * Network.conv(3,3,...) on the input image (network.py:108-132) is ONE taps=1 GEMM with K = k_pad = 32. */
int mv3d_im2col3x3_pad(const float* d_in, int B, int H, int W, int C, int k_pad, void* d_hi, void* d_lo, void* stream);
/* The same first layer as ONE kernel: 3x3 SAME conv + bias + ReLU of a dense float32 (B,H,W,C<=4) input (Network.conv
 * on the image placeholder, network.py:108-132 / MV3D_test.py:51-52), output rendered in `fmt` into the PAD layout
 * (B,H+1,W+1,c_pad), halos zero.  d_w: HWIO (3,3,C,Cout), Cout % 8 == 0.  C <= 3 with Cout == c_pad == 64 runs on the
 * tensor cores (operand rows built in shared memory by the CTA, bf16 hi/lo 3-pass tcgen05 MMAs, TMA tensor stores:
 * error ~2^-17 per product); other shapes use a direct fp32 FMA kernel.  MV3D_SMALL_CIN_MMA=0 forces the latter. */
int mv3d_conv3x3_small_cin(const float* d_in, int B, int H, int W, int C, const float* d_w, const float* d_bias,
                           int Cout, int relu, void* d_out_hi, void* d_out_lo, int c_pad, int fmt, void* stream);
/* PAD -> dense float32 (B,H,W,C) (hi+lo). */
int mv3d_unpad_nhwc(const void* d_hi, const void* d_lo, int B, int H, int W, int C, int c_pad, float* d_out,
                    void* stream);
/* 2x2 stride-2 VALID max-pool on the PAD layout (Network.max_pool, network.py:181-188). */
int mv3d_maxpool2x2_pad(const void* d_in_hi, const void* d_in_lo, int B, int H, int W, int c_pad, void* d_out_hi,
                        void* d_out_lo, void* stream);
/* softmax over channel pairs (2a, 2a+1) of a (rows, 2A) float32 matrix == reshape_layer + softmax +
 * reshape_layer of MV3D_test.py:76-80 (network.py:333-341,399-405); also plain row softmax (A=1). */
int mv3d_softmax_pairs(const float* d_in, int rows, int ld_in, int n_pairs, float* d_out, int ld_out, void* stream);
/* split-K epilogue: out = act(acc + bias) as bf16 hi/lo and/or fp32. */
int mv3d_bias_act(const float* d_acc, int M, int N, int ld_acc, const float* d_bias, int relu, void* d_out_hi,
                  void* d_out_lo, int ld_out, float* d_out_f32, int ld_f32, void* stream);

/* =============================================================================================
 * Training step (BASELINE config 3).  The reference gets its backward pass
 * from TensorFlow autodiff over the graph of lib/networks/MV3D_train.py and the loss/optimizer of
 * lib/fast_rcnn/train_mv.py:94-146; these entry points are that backward pass, hand-written.
 * ============================================================================================= */

/* Backward of Network.max_pool(2,2,2,2,'VALID') (network.py:181-188) fused with the ReLU' of the conv that produced
 * its input: x = pre-pool activation (PAD, HxW), g = gradient of the pooled map (PAD, H/2 x W/2), out PAD HxW. */
int mv3d_maxpool2x2_bwd_pad(const void* d_x_hi, const void* d_x_lo, const void* d_g_hi, const void* d_g_lo, int B,
                            int H, int W, int c_pad, void* d_out_hi, void* d_out_lo, void* stream);
/* db[n] += sum over rows of (g_hi + g_lo)[row, n]  -- bias gradient of conv / fc (network.py:130,395). */
int mv3d_bias_grad(const void* d_g_hi, const void* d_g_lo, long long rows, int ld, int n, float* d_db, void* stream);
/* HWIO / (in,out) float32 weights -> backward-data operand: bf16 hi/lo (cin, taps*cout_pad), taps flipped. */
int mv3d_pack_weights_dgrad(const float* d_w_hwio, int taps, int cin, int cout, int cout_pad, void* d_w_hi,
                            void* d_w_lo, void* stream);
/* dense float32 (B,H,W,C) gradient -> PAD bf16 hi/lo gated by (d_mask_hi > 0) (mask: PAD, same c_pad; may be NULL). */
int mv3d_pad_nhwc_masked(const float* d_in, int B, int H, int W, int C, int c_pad, const void* d_mask_hi, void* d_hi,
                         void* d_lo, void* stream);
/* Network.dropout (network.py:407-409): in place, y = x/keep_prob with probability keep_prob else 0. */
int mv3d_dropout(void* d_hi, void* d_lo, long long rows, int n, int ld, float keep_prob, unsigned long long seed,
                 void* stream);
/* RPN losses + gradients (train_mv.py:94-119, _modified_smooth_l1 :67-84 with sigma).  labels (B,Hf,Wf,A) float32
 * in {-1,0,1}, targets (B,Hf*Wf*A,6), d_counts (B,2) int32 {#label != -1, #label == 1}.  Gradient out: PAD bf16 hi/lo
 * (B,Hf+1,Wf+1,c_pad), channels [0,2A) d rpn_cls_score, [2A,8A) d rpn_bbox_pred.  d_loss[0] += rpn_cross_entropy,
 * d_loss[1] += rpn_loss_box, each the mean over the B frames of the reference's per-frame value. */
int mv3d_rpn_loss(const float* d_cls_score, const float* d_bbox_pred, const float* d_labels, const float* d_targets,
                  const int* d_counts, int B, int Hf, int Wf, int A, int c_pad, float sigma, void* d_grad_hi,
                  void* d_grad_lo, float* d_loss, void* stream);
/* R-CNN losses + gradients (train_mv.py:121-133).  d_rois (R,5) gives each row's frame; d_frame_counts (B) rows per
 * frame.  Gradient out: bf16 hi/lo (R,c_pad): [0,2) d cls_score, [2,2+n_bbox) d bbox_pred.  d_loss[0] += cross_entropy,
 * d_loss[1] += loss_box. */
int mv3d_rcnn_loss(const float* d_cls_score, int ld_cls, const float* d_bbox_pred, int ld_bbox, const int* d_labels,
                   const float* d_targets, int n_bbox, const float* d_rois, const int* d_frame_counts, int B, int R,
                   int c_pad, float sigma, void* d_grad_hi, void* d_grad_lo, float* d_loss, void* stream);
/* tf.train.AdamOptimizer(lr) with TF-1.0 defaults (train_mv.py:144-146) over a flat parameter buffer; grad_scale
 * folds in 1/world_size after the data-parallel gradient all-reduce. */
int mv3d_adam(float* d_theta, const float* d_grad, float* d_m, float* d_v, long long n, float lr, float beta1,
              float beta2, float eps, int step, float grad_scale, void* stream);

/* anchor_target_layer before its random sub-sampling (lib/rpn_msr/anchor_target_layer_tf.py:93-143,164-165).
 *   d_anchors (N,4) int32 all shifted anchors, d_anchors3d (N,6) float64 = bv_anchor_to_lidar(anchors)
 *   d_gt_bv (G,5) / d_gt_3d (G,7) float32.  Outputs: d_max_ov (N) float64 (-1 outside the image), d_argmax (N),
 *   d_code (N) int8: bits 0-1 = label+1, bit 2 = inside, bit 3 = max_overlap < neg_thr; d_targets (N,6) float32
 *   (bbox_transform_3d vs the arg-max GT, zeros outside).  d_gt_max_ws: G x 8 bytes scratch. */
int mv3d_anchor_targets(const int* d_anchors, const double* d_anchors3d, int N, const float* d_gt_bv,
                        const float* d_gt_3d, int G, float im_h, float im_w, double pos_thr, double neg_thr,
                        int clobber, double* d_max_ov, int* d_argmax, unsigned long long* d_gt_max_ws,
                        signed char* d_code, float* d_targets, void* stream);
/* proposal_target_layer_3d, IoU stage (proposal_target_layer_tf.py:38-44,232-236): candidates = rois ++ GT. */
int mv3d_roi_overlaps(const float* d_rois_bv, int R, const float* d_gt_bv, int G, double* d_max_ov, int* d_argmax,
                      void* stream);
/* proposal_target_layer_3d, output stage (:270-298, :80-92): d_keep (K) indexes the candidates, first n_fg are
 * foreground; d_proj = 12 floats (P2.R0).Tr on the DEVICE.  Outputs rois_bv (K,5), rois_img (K,5), labels (K) int32,
 * bbox_targets (K,24*num_classes), rois_3d (K,7); column 0 of the roi blobs = batch_index. */
int mv3d_proposal_targets(const float* d_rois_bv, const float* d_rois_3d, int R, const float* d_gt_bv,
                          const float* d_gt_3d, const float* d_gt_corners, int G, const int* d_keep, int K, int n_fg,
                          const int* d_assign, const float* d_proj, int num_classes, float batch_index,
                          float* d_out_bv, float* d_out_img, int* d_out_labels, float* d_out_targets, float* d_out_3d,
                          void* stream);

/* =============================================================================================
 * Front view (FV).  No reference counterpart: lib/networks/network.py:313-315 returns None for target='fv'.  The
 * semantics are this project's specification (oracle/mv3d_oracle.py: FvGeometry, point_cloud_2_front,
 * lidar_3d_to_fv), after the MV3D paper linked from the reference's README.md:5.
 *   mv3d_fv_raster: points -> (H,W,3) float32 [z, range, reflectance] and/or the PAD bf16 hi/lo trunk input
 *     (H+1,W+1,c_pad); cell = floor((atan2(y,x)-theta_min)/dtheta), floor((phi_max-atan2(z,hypot(x,y)))/dphi) in
 *     float64, x > 0 only, last point in file order wins.  Workspace: mv3d_fv_raster_workspace_bytes(H,W).
 *   mv3d_rois_to_fv: rois_3d (R,7) -> rois_fv (R,5) [batch,col_min,row_min,col_max,row_max] from the 8 corners.
 * ============================================================================================= */
size_t mv3d_fv_raster_workspace_bytes(int H, int W);
int mv3d_fv_raster(const float* d_points, int n_points, int point_stride, int H, int W, double theta_min_rad,
                   double dtheta_rad, double phi_max_rad, double dphi_rad, float* d_top, void* d_pad_hi, void* d_pad_lo,
                   int c_pad, void* d_workspace, size_t workspace_bytes, void* stream);
int mv3d_rois_to_fv(const float* d_rois_3d, int R, const int* d_num_valid, int H, int W, double theta_min_rad,
                    double dtheta_rad, double phi_max_rad, double dphi_rad, float* d_rois_fv, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MV3D_B200_H_ */
