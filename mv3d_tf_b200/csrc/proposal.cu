// proposal_layer_3d on the device.  Replaces lib/rpn_msr/proposal_layer_tf.py:25-202 and the numpy helpers
// it calls (see include/mv3d_b200.h).  Seven small launches, no host round trip:
//   1 decode   : per anchor -- fg score, bbox_transform_inv_3d, lidar_3d_to_bv (numpy float `//` emulated
//                in fp64), clip, 8 corners -> image box (int32, x86 cast semantics), both filters,
//                and a 4096-bin score histogram of the survivors
//   2 scan     : descending exclusive scan of the histogram, M = min(#survivors, pre_nms_topN)
//   3 scatter  : survivors grouped by score bin (64-bit key = orderable(score) << 32 | anchor index)
//   4 rank     : exact rank inside the bin by counting -> descending order, ties higher-index-first
//                (what argsort(kind='stable')[::-1] yields); gathers the sorted BEV boxes
//   5,6 NMS    : nms.cu (mask + single-CTA keep chain, stops at post_nms_topN)
//   7 gather   : blob_bv / blob_img / blob_3d rows
// All float arithmetic mirrors numpy's dtype pipeline; compiled with --fmad=false.
#include "common.cuh"
#include "geom.cuh"

extern "C" size_t mv3d_nms_workspace_bytes(int n_boxes);
extern "C" int mv3d_nms(const float*, int, int, const int*, double, int, int, int*, int*, void*, size_t, void*);

namespace mv3d {

constexpr int kBins = 4096;

struct DecodeConst {
    float M[12];  // (P2.R0).Tr, float32 row-major 3x4
    const float* dM;  // optional: the same 12 floats in DEVICE memory (CUDA-graph replay with a per-frame calib)
    double xn, yn, x_min, y_min, res;
    float clip_x, clip_y;  // im_w - 1, im_h - 1 (float32)
    float min_size;        // RPN_MIN_SIZE * im_scale (float32)
    int img_x_max, img_y_max;  // img_w + 50, img_h + 50
    int Hf, Wf, A, N;
    int ld_prob, ld_deltas;  // floats between consecutive cells of prob / deltas (2A / 6A when dense)
};

__device__ __forceinline__ unsigned int orderable(float s) {
    const unsigned int u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ int score_bin(float s) {
    if (s != s) return kBins - 1;
    const float t = s * (float)kBins;
    if (!(t > 0.f)) return 0;
    return t >= (float)(kBins - 1) ? kBins - 1 : (int)t;
}

struct Decoded {
    float score;
    float p3d[6];
    float bv[4];
    int img[4];
    bool keep;
};

__device__ __forceinline__ Decoded decode_one(const float* __restrict__ prob, const float* __restrict__ deltas,
                                              const float* __restrict__ anchors3d, const DecodeConst& k, int i) {
    Decoded o;
    const int a = i % k.A;
    const int cell = i / k.A;
    o.score = prob[(size_t)cell * k.ld_prob + 2 * a + 1];                    // proposal_layer_tf.py:63
    const float* d = deltas + (size_t)cell * k.ld_deltas + a * 6;              // :105
    const float* an = anchors3d + (size_t)i * 6;
    // bbox_transform_inv_3d (bbox_transform.py:131-136): float32, mul then add
    const float px = __fadd_rn(__fmul_rn(d[0], an[3]), an[0]);
    const float py = __fadd_rn(__fmul_rn(d[1], an[4]), an[1]);
    const float pz = __fadd_rn(__fmul_rn(d[2], an[5]), an[2]);
    const float pl = __fmul_rn((float)exp((double)d[3]), an[3]);
    const float pw = __fmul_rn((float)exp((double)d[4]), an[4]);
    const float ph = __fmul_rn((float)exp((double)d[5]), an[5]);
    o.p3d[0] = px; o.p3d[1] = py; o.p3d[2] = pz; o.p3d[3] = pl; o.p3d[4] = pw; o.p3d[5] = ph;
    // lidar_3d_to_bv (transform.py:132-140) + clip_boxes (bbox_transform.py:178-191): geom.cuh
    const BoxExtents ex = box_extents(px, py, pz, pl, pw, ph);
    BevGrid grid;
    grid.xn = k.xn; grid.yn = k.yn; grid.x_min = k.x_min; grid.y_min = k.y_min; grid.res = k.res;
    grid.clip_x = k.clip_x; grid.clip_y = k.clip_y;
    extents_to_bev_box(grid, ex, o.bv);
    const float x1 = o.bv[0], y1 = o.bv[1], x2 = o.bv[2], y2 = o.bv[3];
    const float xp = ex.xp, xm = ex.xm, yp = ex.yp, ym = ex.ym, zp = ex.zp, zm = ex.zm;
    const float ws = __fadd_rn(__fsub_rn(x2, x1), 1.f), hs = __fadd_rn(__fsub_rn(y2, y1), 1.f);
    const bool keep_size = (ws >= k.min_size) && (hs >= k.min_size);           // _filter_boxes :336-341
    // lidar_3d_to_corners (transform.py:305-313) + lidar_cnr_to_img (:483-500, :369-386)
    float Mloc[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) Mloc[q] = k.dM ? __ldg(k.dM + q) : k.M[q];
    corners_to_img_box(Mloc, xp, xm, yp, ym, zp, zm, o.img);
    const bool keep_img = (-50 <= o.img[0]) && (o.img[2] <= k.img_x_max) && (-50 <= o.img[1]) &&
                          (o.img[3] <= k.img_y_max);                            // _filter_img_boxes :343-352
    o.keep = keep_size && keep_img;
    return o;
}

__global__ void proposal_decode_kernel(const float* __restrict__ prob, const float* __restrict__ deltas,
                                       const float* __restrict__ anchors3d, DecodeConst k, float* __restrict__ score,
                                       float* __restrict__ p3d, float4* __restrict__ pbv, int4* __restrict__ pimg,
                                       unsigned char* __restrict__ keep, int* __restrict__ bin_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k.N) return;
    const Decoded o = decode_one(prob, deltas, anchors3d, k, i);
    score[i] = o.score;
#pragma unroll
    for (int q = 0; q < 6; ++q) p3d[(size_t)i * 6 + q] = o.p3d[q];
    pbv[i] = make_float4(o.bv[0], o.bv[1], o.bv[2], o.bv[3]);
    pimg[i] = make_int4(o.img[0], o.img[1], o.img[2], o.img[3]);
    keep[i] = o.keep ? 1 : 0;
    if (o.keep && bin_count) atomicAdd(&bin_count[score_bin(o.score)], 1);
}

// offsets in DESCENDING bin order; meta[0] = #survivors, meta[1] = M = min(#survivors, pre_top_n)
__global__ void proposal_scan_kernel(const int* __restrict__ bin_count, int* __restrict__ bin_offset,
                                     int* __restrict__ bin_cursor, int pre_top_n, int* __restrict__ meta) {
    __shared__ int part[1024];
    const int t = threadIdx.x;  // 1024 threads, 4 bins each; thread 0 owns the HIGHEST bins
    int c[4], s = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { c[q] = bin_count[kBins - 1 - (t * 4 + q)]; s += c[q]; }
    part[t] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const int v = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int run = part[t] - s;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int b = kBins - 1 - (t * 4 + q);
        bin_offset[b] = run;
        bin_cursor[b] = 0;
        run += c[q];
    }
    if (t == 1023) {
        const int total = part[1023];
        meta[0] = total;
        meta[1] = (pre_top_n > 0 && total > pre_top_n) ? pre_top_n : total;
    }
}

__global__ void proposal_scatter_kernel(const float* __restrict__ score, const unsigned char* __restrict__ keep, int N,
                                        const int* __restrict__ bin_offset, int* __restrict__ bin_cursor,
                                        unsigned long long* __restrict__ grouped) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || !keep[i]) return;
    const float s = score[i];
    const int b = score_bin(s);
    const int pos = bin_offset[b] + atomicAdd(&bin_cursor[b], 1);
    grouped[pos] = ((unsigned long long)orderable(s) << 32) | (unsigned int)i;
}

// Exact rank inside the score bin = position in the order the reference's NMS sees.  The reference sorts twice:
// `scores.ravel().argsort()[::-1][:pre_nms_topN]` (proposal_layer_tf.py:161-163), then again inside nms()
// (cpu_nms.pyx:25 / gpu_nms.pyx: `scores.argsort()[::-1]` on the already-descending array).  numpy's default sort leaves
// tie order unspecified; the pinned rule (oracle.argsort_desc) is "stable ascending, reversed", under which
// the first sort puts ties higher-index-first and the second one REVERSES every tie group that survived the top-N cut.
// Both are applied here: r = #greater + #(equal, higher index) is the first sort's position; a box past the cut is
// dropped; the tie group [#greater, min(#greater + #equal, M)) is then mirrored.
// Cost: one pass over the bin per element (sum of bin_count^2 compares).  Scores concentrated in one bin (a freshly
// initialised or a saturated RPN: ~35 k survivors in one bin) make that ~1e9 compares spread over ~35 k resident
// threads, every load a warp broadcast: ~0.2 ms worst case instead of ~10 us -- bounded, so no second code path.
__global__ void proposal_rank_kernel(const unsigned long long* __restrict__ grouped, const float* __restrict__ score,
                                     const int* __restrict__ bin_count, const int* __restrict__ bin_offset,
                                     const int* __restrict__ meta, const float4* __restrict__ pbv,
                                     int* __restrict__ sorted_idx, float4* __restrict__ sorted_box) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= meta[0]) return;
    const unsigned long long key = grouped[pos];
    const int idx = (int)(key & 0xffffffffu);
    const unsigned int skey = (unsigned int)(key >> 32);
    const int b = score_bin(score[idx]);
    const int beg = bin_offset[b];
    const int M = meta[1];
    if (beg >= M) return;  // the whole bin ranks past pre_nms_topN
    const int end = beg + bin_count[b];
    int greater = beg, equal_hi = 0, equal = 0;
    for (int q = beg; q < end; ++q) {
        const unsigned long long other = grouped[q];
        const unsigned int okey = (unsigned int)(other >> 32);
        greater += (okey > skey) ? 1 : 0;
        equal += (okey == skey) ? 1 : 0;
        equal_hi += (okey == skey && other > key) ? 1 : 0;
    }
    const int first = greater + equal_hi;            // position after the first sort
    if (first < M) {
        const int last = min(greater + equal, M) - 1;  // the tie group's last position inside the cut
        const int rank = greater + (last - first);      // mirrored by the second sort
        sorted_idx[rank] = idx;
        sorted_box[rank] = pbv[idx];
    }
}

__global__ void proposal_gather_kernel(const int* __restrict__ keep_list, const int* __restrict__ num_keep,
                                       const int* __restrict__ sorted_idx, const float* __restrict__ score,
                                       const float* __restrict__ p3d, const float4* __restrict__ pbv,
                                       const int4* __restrict__ pimg, int cap, float batch_index,
                                       float* __restrict__ blob_bv, float* __restrict__ blob_img,
                                       float* __restrict__ blob_3d, float* __restrict__ out_score,
                                       int* __restrict__ out_anchor, int* __restrict__ num_out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0 && num_out) *num_out = *num_keep;
    if (r >= cap) return;
    const bool valid = r < *num_keep;
    int idx = -1;
    float4 bv = make_float4(0, 0, 0, 0);
    int4 im = make_int4(0, 0, 0, 0);
    float sc = 0.f, p[6] = {0, 0, 0, 0, 0, 0};
    if (valid) {
        idx = sorted_idx[keep_list[r]];
        bv = pbv[idx];
        im = pimg[idx];
        sc = score[idx];
#pragma unroll
        for (int q = 0; q < 6; ++q) p[q] = p3d[(size_t)idx * 6 + q];
    }
    const float bi = valid ? batch_index : 0.f;
    float* o = blob_bv + (size_t)r * 5;
    o[0] = bi; o[1] = bv.x; o[2] = bv.y; o[3] = bv.z; o[4] = bv.w;
    o = blob_img + (size_t)r * 5;
    o[0] = bi; o[1] = (float)im.x; o[2] = (float)im.y; o[3] = (float)im.z; o[4] = (float)im.w;
    o = blob_3d + (size_t)r * 7;
    o[0] = bi;
#pragma unroll
    for (int q = 0; q < 6; ++q) o[1 + q] = p[q];
    if (out_score) out_score[r] = sc;
    if (out_anchor) out_anchor[r] = idx;
}

struct ProposalWs {
    size_t score, p3d, pbv, pimg, keep, bin_count, bin_offset, bin_cursor, meta, grouped, sorted_idx, sorted_box,
        keep_list, num_keep, nms, total;
};

static ProposalWs proposal_layout(const mv3d_proposal_params* p) {
    ProposalWs w;
    const size_t N = (size_t)p->Hf * p->Wf * p->A;
    const size_t cap = (p->pre_nms_top_n > 0 && (size_t)p->pre_nms_top_n < N) ? (size_t)p->pre_nms_top_n : N;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
    w.score = take(4 * N); w.p3d = take(24 * N); w.pbv = take(16 * N); w.pimg = take(16 * N); w.keep = take(N);
    w.bin_count = take(4 * kBins); w.bin_offset = take(4 * kBins); w.bin_cursor = take(4 * kBins); w.meta = take(16);
    w.grouped = take(8 * N); w.sorted_idx = take(4 * cap); w.sorted_box = take(16 * cap);
    w.keep_list = take(4 * cap); w.num_keep = take(16);
    w.nms = take(mv3d_nms_workspace_bytes((int)cap));
    w.total = o;
    return w;
}

static int make_decode_const(const mv3d_proposal_params* p, const float* h_proj, DecodeConst* k) {
    if (!p || (!h_proj && !p->d_proj) || p->Hf <= 0 || p->Wf <= 0 || p->A <= 0) return MV3D_ERR_ARG;
    for (int i = 0; i < 12; ++i) k->M[i] = h_proj ? h_proj[i] : 0.f;
    k->dM = p->d_proj;
    k->xn = p->xn; k->yn = p->yn; k->x_min = p->x_min; k->y_min = p->y_min; k->res = p->res;
    k->clip_x = p->im_w - 1.0f;  // im_shape[1] - 1 on a float32 array element
    k->clip_y = p->im_h - 1.0f;
    k->min_size = p->min_size * p->im_scale;
    k->img_x_max = (int)p->img_w + 50;
    k->img_y_max = (int)p->img_h + 50;
    k->Hf = p->Hf; k->Wf = p->Wf; k->A = p->A; k->N = p->Hf * p->Wf * p->A;
    k->ld_prob = p->ld_prob > 0 ? p->ld_prob : 2 * p->A;
    k->ld_deltas = p->ld_deltas > 0 ? p->ld_deltas : 6 * p->A;
    if (k->ld_prob < 2 * p->A || k->ld_deltas < 6 * p->A) return MV3D_ERR_ARG;
    return MV3D_OK;
}

}  // namespace mv3d

using namespace mv3d;

extern "C" __attribute__((visibility("default"))) size_t mv3d_proposal_workspace_bytes(const mv3d_proposal_params* p) {
    if (!p) return 0;
    return proposal_layout(p).total;
}

extern "C" __attribute__((visibility("default"))) int mv3d_proposal_decode(
    const float* d_prob, const float* d_deltas, const float* d_anchors3d, const float* h_proj,
    const mv3d_proposal_params* p, float* d_score, float* d_p3d, float* d_pbv, int* d_pimg, unsigned char* d_keep,
    void* stream) {
    DecodeConst k;
    int rc = make_decode_const(p, h_proj, &k);
    if (rc != MV3D_OK) return rc;
    MV3D_REQUIRE(d_prob && d_deltas && d_anchors3d && d_score && d_p3d && d_pbv && d_pimg && d_keep);
    proposal_decode_kernel<<<ceil_div(k.N, 128), 128, 0, (cudaStream_t)stream>>>(
        d_prob, d_deltas, d_anchors3d, k, d_score, d_p3d, (float4*)d_pbv, (int4*)d_pimg, d_keep, nullptr);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

extern "C" __attribute__((visibility("default"))) int mv3d_proposal_layer_3d(
    const float* d_prob, const float* d_deltas, const float* d_anchors3d, const float* h_proj,
    const mv3d_proposal_params* p, float* d_blob_bv, float* d_blob_img, float* d_blob_3d, float* d_scores,
    int* d_anchor_index, int* d_num_out, void* d_workspace, size_t workspace_bytes, void* stream) {
    DecodeConst k;
    int rc = make_decode_const(p, h_proj, &k);
    if (rc != MV3D_OK) return rc;
    MV3D_REQUIRE(d_prob && d_deltas && d_anchors3d && d_blob_bv && d_blob_img && d_blob_3d && d_num_out);
    const ProposalWs w = proposal_layout(p);
    if (!d_workspace || workspace_bytes < w.total) return MV3D_ERR_WORKSPACE;
    char* ws = static_cast<char*>(d_workspace);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int N = k.N;
    const int cap = (p->pre_nms_top_n > 0 && p->pre_nms_top_n < N) ? p->pre_nms_top_n : N;
    const int out_cap = (p->post_nms_top_n > 0 && p->post_nms_top_n < cap) ? p->post_nms_top_n : cap;
    float* score = (float*)(ws + w.score);
    float* p3d = (float*)(ws + w.p3d);
    float4* pbv = (float4*)(ws + w.pbv);
    int4* pimg = (int4*)(ws + w.pimg);
    unsigned char* keep = (unsigned char*)(ws + w.keep);
    int* bin_count = (int*)(ws + w.bin_count);
    int* bin_offset = (int*)(ws + w.bin_offset);
    int* bin_cursor = (int*)(ws + w.bin_cursor);
    int* meta = (int*)(ws + w.meta);
    unsigned long long* grouped = (unsigned long long*)(ws + w.grouped);
    int* sorted_idx = (int*)(ws + w.sorted_idx);
    float4* sorted_box = (float4*)(ws + w.sorted_box);
    int* keep_list = (int*)(ws + w.keep_list);
    int* num_keep = (int*)(ws + w.num_keep);

    cudaError_t e = cudaMemsetAsync(bin_count, 0, sizeof(int) * kBins, s);
    if (e != cudaSuccess) { set_last_cuda_error(e); return MV3D_ERR_LAUNCH; }
    proposal_decode_kernel<<<ceil_div(N, 128), 128, 0, s>>>(d_prob, d_deltas, d_anchors3d, k, score, p3d, pbv, pimg,
                                                             keep, bin_count);
    proposal_scan_kernel<<<1, 1024, 0, s>>>(bin_count, bin_offset, bin_cursor, p->pre_nms_top_n, meta);
    proposal_scatter_kernel<<<ceil_div(N, 256), 256, 0, s>>>(score, keep, N, bin_offset, bin_cursor, grouped);
    proposal_rank_kernel<<<ceil_div(N, 128), 128, 0, s>>>(grouped, score, bin_count, bin_offset, meta, pbv, sorted_idx,
                                                           sorted_box);
    MV3D_CHECK_LAUNCH();
    rc = mv3d_nms((const float*)sorted_box, cap, 4, meta + 1, p->nms_thresh, p->nms_rule_ge, out_cap, keep_list,
                  num_keep, ws + w.nms, w.total - w.nms, s);
    if (rc != MV3D_OK) return rc;
    proposal_gather_kernel<<<ceil_div(out_cap, 128), 128, 0, s>>>(keep_list, num_keep, sorted_idx, score, p3d, pbv, pimg,
                                                                  out_cap, p->batch_index, d_blob_bv, d_blob_img,
                                                                  d_blob_3d, d_scores, d_anchor_index, d_num_out);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}
