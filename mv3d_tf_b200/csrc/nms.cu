// Greedy NMS entirely on the device.  Replaces lib/nms/cpu_nms.pyx:17-68 (rule '>=' in double) and
// lib/nms/nms_kernel.cu:34-144 (rule '>' in float; there the keep-chain is reduced on the HOST after an
// 18 MB mask D2H -- here it never leaves the GPU).
//
//  nms_mask_kernel  : one 256-thread CTA per UPPER-TRIANGULAR 64x64 tile (1-D grid over the T(T+1)/2 tiles, no
//                     empty CTAs).  The 64 row boxes sit in shared memory; each lane keeps two column boxes (+ areas)
//                     in registers (128-bit loads), each warp walks 8 rows: the row box is a shared-memory broadcast,
//                     every lane evaluates its two IoUs and two __ballot_sync form the row's 64-bit suppression word.
//  nms_reduce_kernel: ONE CTA walks the keep chain in super-blocks of S = 1024 candidates.  Per super-block ONE round
//                     of independent global loads: (a) the `removed` bits of its candidates = OR of the mask rows of
//                     every box kept so far (gathered, not scattered), (b) its S x S/64 diagonal tile into shared
//                     memory.  Warp 0 then resolves the super-block with one step per KEPT box (not per candidate):
//                     ballot over the 16 word lanes -> first un-suppressed candidate -> OR its diagonal row (one
//                     shared-memory word per lane).  Stops as soon as max_keep survivors exist (the reference computes
//                     all survivors and slices [:post_nms_topN]; the prefix is identical).
//
// IoU arithmetic is the reference's, float32 with IEEE roundings (file compiled with --fmad=false):
//   area = (x2-x1+1)*(y2-y1+1); w = max(0, min(x2)-max(x1)+1); ovr = w*h / (area_i + area_j - w*h).
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace mv3d {

using namespace ptx;

constexpr int kNmsTile = 64;
constexpr int kMaskThreads = 256;

__device__ __forceinline__ float4 load_box(const float* boxes, int stride, int i) {
    if (stride == 4) return __ldg(reinterpret_cast<const float4*>(boxes) + i);
    const float* p = boxes + (size_t)i * stride;
    return make_float4(p[0], p[1], p[2], p[3]);
}

// thresh_ge = the smallest float32 t with (double)t >= thresh, so that `(double)ovr >= thresh` (cpu_nms.pyx:65) is the
// single float compare `ovr >= t`; thresh_gt = (float)thresh for the CUDA rule `ovr > 0.7f` (nms_kernel.cu:31,71).
// `positive` (thresh > 0, uniform): boxes that do not intersect have inter = 0 and ovr = +-0 or NaN, which never reaches
// a positive threshold -- the division (most of the work: random pairs rarely intersect) is skipped for them.
__device__ __forceinline__ bool suppresses(const float4 a, const float a_area, const float4 b, const float b_area,
                                           const int rule_ge, const float thresh_ge, const float thresh_gt,
                                           const bool positive) {
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.f, xx2 - xx1 + 1.f), h = fmaxf(0.f, yy2 - yy1 + 1.f);
    if (positive && !(w > 0.f && h > 0.f)) return false;
    const float inter = w * h;
    const float ovr = inter / (a_area + b_area - inter);
    return rule_ge ? (ovr >= thresh_ge) : (ovr > thresh_gt);
}

__global__ void __launch_bounds__(kMaskThreads)
nms_mask_kernel(const float* __restrict__ boxes, int n_max, int stride, const int* __restrict__ d_n, double thresh,
                int rule_ge, int nwords, unsigned long long* __restrict__ mask) {
    // tile id -> (row_blk, col_blk >= row_blk): row r starts at r*T - r(r-1)/2
    const int T = nwords;
    const int t = blockIdx.x;
    int row_blk = (int)((2.0f * T + 1.0f - sqrtf((2.0f * T + 1.0f) * (2.0f * T + 1.0f) - 8.0f * (float)t)) * 0.5f);
    row_blk = max(0, min(row_blk, T - 1));
    while (row_blk > 0 && row_blk * T - row_blk * (row_blk - 1) / 2 > t) --row_blk;
    while ((row_blk + 1) * T - (row_blk + 1) * row_blk / 2 <= t) ++row_blk;
    const int col_blk = row_blk + (t - (row_blk * T - row_blk * (row_blk - 1) / 2));
    int n = n_max;
    if (d_n) n = min(n, *d_n);
    if (col_blk * kNmsTile >= n) return;   // (row_blk <= col_blk)
    __shared__ float4 rbox[kNmsTile];
    __shared__ float rarea[kNmsTile];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < kNmsTile) {
        const int i = row_blk * kNmsTile + tid;
        const float4 a = i < n ? load_box(boxes, stride, i) : make_float4(0.f, 0.f, 0.f, 0.f);
        rbox[tid] = a;
        rarea[tid] = (a.z - a.x + 1.f) * (a.w - a.y + 1.f);
    }
    const int c0 = col_blk * kNmsTile + lane, c1 = c0 + 32;
    const bool v0 = c0 < n, v1 = c1 < n;
    const float4 b0 = v0 ? load_box(boxes, stride, c0) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 b1 = v1 ? load_box(boxes, stride, c1) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float area0 = (b0.z - b0.x + 1.f) * (b0.w - b0.y + 1.f), area1 = (b1.z - b1.x + 1.f) * (b1.w - b1.y + 1.f);
    const float thresh_gt = (float)thresh;
    float thresh_ge = (float)thresh;
    if ((double)thresh_ge < thresh) thresh_ge = nextafterf(thresh_ge, INFINITY);
    const bool positive = thresh > 0.0;
    const bool diagonal = row_blk == col_blk;
    __syncthreads();
    unsigned long long mine = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int rt = warp * 8 + r;                      // row inside the tile
        const float4 a = rbox[rt];
        const float a_area = rarea[rt];
        bool s0 = v0 && suppresses(a, a_area, b0, area0, rule_ge, thresh_ge, thresh_gt, positive);
        bool s1 = v1 && suppresses(a, a_area, b1, area1, rule_ge, thresh_ge, thresh_gt, positive);
        if (diagonal) { s0 = s0 && lane > rt; s1 = s1 && lane + 32 > rt; }   // only later boxes can be suppressed
        const unsigned lo = __ballot_sync(0xffffffffu, s0), hi = __ballot_sync(0xffffffffu, s1);
        if (lane == r) mine = ((unsigned long long)hi << 32) | lo;
    }
    const int i = row_blk * kNmsTile + warp * 8 + lane;
    if (lane < 8 && i < n) mask[(size_t)i * nwords + col_blk] = mine;
}

constexpr int kReduceThreads = 1024;
constexpr int kMaxWords = 1024;      // up to 65536 boxes
constexpr int kSuper = 1024;         // candidates per super-block of the keep chain
constexpr int kSuperWords = kSuper / kNmsTile;   // 16: one word lane each in the resolving warp
constexpr int kReduceSmem = kSuper * kSuperWords * (int)sizeof(unsigned long long);   // 128 KB diagonal tile

__global__ void __launch_bounds__(kReduceThreads)
nms_reduce_kernel(const unsigned long long* __restrict__ mask, int n_max, const int* __restrict__ d_n, int nwords,
                  int max_keep, int* __restrict__ keep_out, int* __restrict__ num_out) {
    extern __shared__ unsigned long long diag[];               // [kSuper][kSuperWords]
    __shared__ unsigned long long cur_s[kSuperWords];
    __shared__ int kept_s[kSuper];                             // kept candidates of the current super-block
    __shared__ int count_s;
    int n = n_max;
    if (d_n) n = min(n, *d_n);
    if (max_keep <= 0 || max_keep > n) max_keep = n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) count_s = 0;
    const int nsb = (n + kSuper - 1) / kSuper;
    for (int B = 0; B < nsb; ++B) {
        const int base = B * kSuper, w0 = B * kSuperWords;
        const int nw = min(kSuperWords, nwords - w0);           // words of this super-block
        const int rows = min(kSuper, n - base);
        if (tid < kSuperWords) cur_s[tid] = 0;
        __syncthreads();                                        // count_s / keep_out of the previous super-block visible
        const int cnt = count_s;
        {   // (a) removed bits of this super-block's candidates: OR of the rows of every box kept so far.  One LANE per
            //     kept box: its index, then its 16 words -- independent loads, two round trips in all -- and the warp
            //     combines each word with two 32-bit OR-reductions (the kept list is at most a few thousand long).
            for (int k0 = warp * 32; k0 < cnt; k0 += kReduceThreads) {
                const int k = k0 + lane;
                const unsigned long long* row = (k < cnt) ? mask + (size_t)keep_out[k] * nwords + w0 : nullptr;
                unsigned long long v[kSuperWords];
#pragma unroll
                for (int w = 0; w < kSuperWords; ++w) v[w] = (row != nullptr && w < nw) ? __ldg(row + w) : 0ull;
#pragma unroll
                for (int w = 0; w < kSuperWords; ++w) {
                    const unsigned lo32 = __reduce_or_sync(0xffffffffu, (unsigned)v[w]);
                    const unsigned hi32 = __reduce_or_sync(0xffffffffu, (unsigned)(v[w] >> 32));
                    if (lane == w && (lo32 | hi32)) atomicOr(&cur_s[w], ((unsigned long long)hi32 << 32) | lo32);
                }
            }
            // (b) the diagonal tile; words left of a row's own 64-block were never written (upper-triangular mask)
            for (int e = tid; e < rows * kSuperWords; e += kReduceThreads) {
                const int r = e >> 4, ww = e & (kSuperWords - 1);
                diag[e] = (ww < nw && ww >= (r >> 6)) ? __ldg(mask + (size_t)(base + r) * nwords + w0 + ww) : 0ull;
            }
        }
        __syncthreads();
        if (warp == 0) {
            // Resolve with the whole warp in lock-step (uniform registers, no ballots on the chain): inside one 64-candidate
            // word the chain is ffs -> one broadcast shared-memory load of the kept row's word -> OR (~45 cycles per KEPT
            // box); moving on to the next word, the lanes OR the rows of this super-block's kept boxes for that word in
            // parallel (two 32-bit warp OR-reductions).
            int c = cnt, m = 0;                                      // kept so far overall / in this super-block
            for (int w = 0; w < nw && c < max_keep; ++w) {
                unsigned long long acc = 0;                          // suppression of word w by this super-block's kept boxes
                for (int i = lane; i < m; i += 32) acc |= diag[kept_s[i] * kSuperWords + w];
                const unsigned lo32 = __reduce_or_sync(0xffffffffu, (unsigned)acc);
                const unsigned hi32 = __reduce_or_sync(0xffffffffu, (unsigned)(acc >> 32));
                unsigned long long cur = cur_s[w] | ((unsigned long long)hi32 << 32) | lo32;
                const int valid = rows - w * kNmsTile;               // candidates of this word that exist
                if (valid < kNmsTile) cur |= valid <= 0 ? ~0ull : (~0ull << valid);
                while (c < max_keep) {
                    const unsigned long long avail = ~cur;           // not suppressed, not yet visited
                    if (avail == 0ull) break;
                    const int bit = __ffsll((long long)avail) - 1;
                    const int k = w * kNmsTile + bit;
                    if (lane == 0) { keep_out[c] = base + k; kept_s[m] = k; }
                    ++c; ++m;
                    cur |= diag[k * kSuperWords + w] | (1ull << bit);
                }
                __syncwarp();                                        // kept_s of this word visible to all lanes
            }
            if (lane == 0) count_s = c;
        }
        __syncthreads();
        if (count_s >= max_keep) break;
    }
    __syncthreads();
    if (tid == 0) *num_out = count_s;
}

// ----------------------------------------------------------------------------------------------------------------
// nms_lazy_kernel: greedy NMS that only ever tests a candidate against KEPT boxes (and its own block), for callers
// that want at most max_keep <= 2048 survivors (the proposal layer: 300 at test time, 2000 in training).  The all-pairs
// mask above costs n^2/2 tests (18 M at n = 6000) of which the keep chain reads the rows of the <= max_keep kept boxes
// only, and it stops at max_keep -- here the work is (candidates scanned) x (kept so far).
// One cluster of 8 CTAs x 512 threads walks the sorted candidates in blocks of 512:
//   (a) every CTA tests the block's 512 candidates (one per thread) against its 1/8 slice of the kept list (boxes in
//       shared memory, broadcast reads) and ORs the 512 suppression bits into ALL eight CTAs' shared memory (DSMEM atomics);
//   (b) every CTA forms 1/8 of the block's own 512 x 512 upper-triangular mask (a warp per row, __ballot_sync words)
//       and stores its rows into all eight CTAs' shared memory;
//   one cluster barrier; then EVERY CTA resolves the block redundantly (same inputs, same result: no second exchange):
//   if the block's own mask is empty, all un-suppressed candidates are kept at once (prefix popcount); otherwise warp 0
//   walks the chain with one step per KEPT box (ffs -> one shared-memory word -> OR), the later words' suppression
//   accumulating in lanes 0..7 off the critical path.  Exchange buffers are double-buffered by block parity, so one
//   barrier per block suffices.  Same predicate, same order => the same survivor list as the mask + reduce path.
// ----------------------------------------------------------------------------------------------------------------
constexpr int kLazyC = 8;            // CTAs per cluster
constexpr int kLazyBS = 512;         // candidates per block
constexpr int kLazyThreads = 512;
constexpr int kLazyWords = kLazyBS / 64;
constexpr int kLazyMaxKeep = 2048;

struct LazySmem {
    float4 kept_box[kLazyMaxKeep];
    float4 cand_box[kLazyBS];
    unsigned long long rowsT[2][kLazyWords][kLazyBS];   // [parity][word][row]: word-major, conflict-free for the chain
    unsigned long long flags[2][kLazyWords + 1];        // suppressed-by-kept bits; [8] != 0: the block's own mask has a bit
    float kept_area[kLazyMaxKeep];
    float cand_area[kLazyBS];
    int kept_blk[kLazyBS];
    int m_s;
};

__global__ void __cluster_dims__(kLazyC, 1, 1) __launch_bounds__(kLazyThreads)
nms_lazy_kernel(const float* __restrict__ boxes, int n_max, int stride, const int* __restrict__ d_n, double thresh,
                int rule_ge, int max_keep, int* __restrict__ keep_out, int* __restrict__ num_out) {
    extern __shared__ __align__(16) unsigned char lazy_smem_raw[];
    LazySmem& S = *reinterpret_cast<LazySmem*>(lazy_smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int n = n_max;
    if (d_n) n = min(n, *d_n);
    if (max_keep <= 0 || max_keep > n) max_keep = n;
    const float thresh_gt = (float)thresh;
    float thresh_ge = (float)thresh;
    if ((double)thresh_ge < thresh) thresh_ge = nextafterf(thresh_ge, INFINITY);
    const bool positive = thresh > 0.0;
    if (tid < 2 * (kLazyWords + 1)) (&S.flags[0][0])[tid] = 0ull;
    cluster.sync();                       // every CTA's exchange buffers exist and are clear before the first remote write
    int cnt = 0;
    for (int b = 0, base = 0; base < n && cnt < max_keep; ++b, base += kLazyBS) {
        const int p = b & 1;
        const int valid = min(kLazyBS, n - base);
        const float4 box = tid < valid ? load_box(boxes, stride, base + tid) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float area = (box.z - box.x + 1.f) * (box.w - box.y + 1.f);
        S.cand_box[tid] = box;
        S.cand_area[tid] = area;
        // (a) against this CTA's slice of the kept list
        bool sup = false;
        if (tid < valid)
            for (int i = rank; i < cnt; i += kLazyC)
                if (suppresses(S.kept_box[i], S.kept_area[i], box, area, rule_ge, thresh_ge, thresh_gt, positive)) { sup = true; break; }
        const unsigned bal = __ballot_sync(0xffffffffu, sup);
        if (lane == 0 && bal != 0u) {   // this warp's 32 candidates = one 32-bit half of a flags word
            const uint32_t a = smem_u32(&S.flags[p][warp >> 1]) + 4u * (warp & 1);
            for (int r = 0; r < kLazyC; ++r) red_or_cluster_u32(mapa_u32(a, r), bal);
        }
        __syncthreads();                  // cand_box / cand_area complete
        // (b) rows rank, rank + 8, ... of the block's own mask
        bool any = false;
        for (int q = warp; q < kLazyBS / kLazyC; q += kLazyThreads / 32) {
            const int r = rank + kLazyC * q;
            unsigned long long mine = 0ull;
            if (r < valid) {
                const float4 a = S.cand_box[r];
                const float a_area = S.cand_area[r];
                for (int ch = r >> 5; ch < ((valid + 31) >> 5); ++ch) {
                    const int col = ch * 32 + lane;
                    const bool s = col > r && col < valid &&
                                   suppresses(a, a_area, S.cand_box[col], S.cand_area[col], rule_ge, thresh_ge, thresh_gt, positive);
                    const unsigned w32 = __ballot_sync(0xffffffffu, s);
                    if (lane == (ch >> 1)) mine |= (unsigned long long)w32 << (32 * (ch & 1));
                }
            }
            any |= mine != 0ull;
            if (lane < kLazyWords) {
                const uint32_t a = smem_u32(&S.rowsT[p][lane][r]);
                for (int rr = 0; rr < kLazyC; ++rr) st_cluster_u64(mapa_u32(a, rr), mine);
            }
        }
        if (__any_sync(0xffffffffu, any) && lane == 0)
            for (int rr = 0; rr < kLazyC; ++rr) red_or_cluster_u32(mapa_u32(smem_u32(&S.flags[p][kLazyWords]), rr), 1u);
        cluster.sync();                   // flags[p] and rowsT[p] of this block complete in every CTA
        int m;
        if (S.flags[p][kLazyWords] == 0ull) {
            // nobody in the block suppresses anybody in it: every candidate the kept list left alone survives
            int before = 0, total = 0;
            bool mine_ok = false;
#pragma unroll
            for (int w = 0; w < kLazyWords; ++w) {
                unsigned long long av = ~S.flags[p][w];
                const int vb = valid - w * 64;
                av = vb <= 0 ? 0ull : (vb < 64 ? av & ((1ull << vb) - 1ull) : av);
                total += __popcll(av);
                if (w < (tid >> 6)) before += __popcll(av);
                if (w == (tid >> 6)) {
                    before += __popcll(av & ((1ull << (tid & 63)) - 1ull));
                    mine_ok = (av >> (tid & 63)) & 1ull;
                }
            }
            m = min(total, max_keep - cnt);
            if (mine_ok && before < m) S.kept_blk[before] = tid;
        } else {
            if (warp == 0) {
                int c = cnt, mm = 0;
                unsigned long long pend = 0ull;     // lane l < 8: suppression of word l by this block's kept boxes so far
                const int nw = (valid + 63) >> 6;
                for (int w = 0; w < nw && c < max_keep; ++w) {
                    unsigned long long cur = S.flags[p][w] | __shfl_sync(0xffffffffu, pend, w);
                    const int vb = valid - w * 64;
                    if (vb < 64) cur |= ~0ull << vb;
                    while (c < max_keep) {
                        const unsigned long long avail = ~cur;
                        if (avail == 0ull) break;
                        const int bit = __ffsll((long long)avail) - 1;
                        const int k = w * 64 + bit;
                        if (lane == 0) S.kept_blk[mm] = k;
                        pend |= S.rowsT[p][lane & (kLazyWords - 1)][k];
                        cur |= S.rowsT[p][w][k] | (1ull << bit);
                        ++c; ++mm;
                    }
                }
                if (lane == 0) S.m_s = mm;
            }
            __syncthreads();
            m = S.m_s;
        }
        __syncthreads();                  // kept_blk complete
        for (int i = tid; i < m; i += kLazyThreads) {
            const int k = S.kept_blk[i];
            S.kept_box[cnt + i] = S.cand_box[k];
            S.kept_area[cnt + i] = S.cand_area[k];
            if (rank == 0) keep_out[cnt + i] = base + k;
        }
        if (tid <= kLazyWords) S.flags[p][tid] = 0ull;   // clear for block b + 2 (peers write it only after the next barrier)
        cnt += m;
        __syncthreads();                  // kept list extended; cand_box free for the next block
    }
    if (rank == 0 && tid == 0) *num_out = cnt;
    cluster.sync();                       // no CTA leaves while a peer may still address its shared memory
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
    static int lazy = -1;   // MV3D_NMS_LAZY=0: always the all-pairs mask + reduce path (A/B comparisons)
    if (lazy < 0) { const char* e = getenv("MV3D_NMS_LAZY"); lazy = e ? atoi(e) : 1; }
    if (lazy && max_keep > 0 && max_keep <= kLazyMaxKeep) {
        static bool lazy_attr = false;
        if (!lazy_attr) {
            cudaError_t e = cudaFuncSetAttribute(nms_lazy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LazySmem));
            if (e != cudaSuccess) { set_last_cuda_error(e); return MV3D_ERR_LAUNCH; }
            lazy_attr = true;
        }
        nms_lazy_kernel<<<kLazyC, kLazyThreads, sizeof(LazySmem), s>>>(d_boxes, n_boxes, box_stride, d_n_boxes, thresh, rule_ge,
                                                                      max_keep, d_keep_out, d_num_out);
        MV3D_CHECK_LAUNCH();
        return MV3D_OK;
    }
    const int nwords = (int)nms_words(n_boxes);
    MV3D_REQUIRE(nwords <= kMaxWords);
    if (!d_workspace || workspace_bytes < mv3d_nms_workspace_bytes(n_boxes)) return MV3D_ERR_WORKSPACE;
    unsigned long long* mask = static_cast<unsigned long long*>(d_workspace);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(nms_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kReduceSmem);
        if (e != cudaSuccess) { set_last_cuda_error(e); return MV3D_ERR_LAUNCH; }
        attr_set = true;
    }
    const int tiles = nwords * (nwords + 1) / 2;
    nms_mask_kernel<<<tiles, kMaskThreads, 0, s>>>(d_boxes, n_boxes, box_stride, d_n_boxes, thresh, rule_ge, nwords, mask);
    nms_reduce_kernel<<<1, kReduceThreads, kReduceSmem, s>>>(mask, n_boxes, d_n_boxes, nwords, max_keep, d_keep_out,
                                                             d_num_out);
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
