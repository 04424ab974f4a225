// Weight gradient of conv3x3 / 1x1 / fc as a tcgen05 GEMM whose reduction dimension is the PIXEL (row) index:
//     dW[t, c, n] += sum_p X[p + shift_t, c] * G[p, n]          (t = tap, c = input channel, n = output channel)
// This is the backward-filter pass TensorFlow derives for tf.nn.conv2d / xw_plus_b in the reference's train graph
// (lib/fast_rcnn/train_mv.py:146 `AdamOptimizer(lr).minimize(loss)` over lib/networks/network.py:114,395).
//
// Both operands are stored row = pixel, channels contiguous (the PAD activation layout), so K (pixels) is the STRIDED
// dimension of both: they are fed to the tensor core as MN-major operands (instruction-descriptor bits 15/16) straight
// from the TMA boxes -- no transposition pass.  A = G (M = 128 output channels = two 64-channel swizzle atoms),
// B = X window (N = BNX input channels), D[n, c] accumulates in TMEM, one column block per tap.  The three kw taps
// of one kernel row read the same X rows displaced by one pixel, so ONE (KC+8)-row window box serves three taps
// (descriptor start advanced by kw rows), exactly like the forward tap-reuse kernel.
// Work item = (row split, tap group, 128-channel tile of G, BNX-channel tile of X); partial sums of the row splits
// are combined with red.global.add.f32 into dW, laid out (taps, cin, cout) = the reference's HWIO.
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace mv3d {

using namespace ptx;

constexpr int kWgThreads = 192;
constexpr int kWgKC = 64;           // pixels (K) per pipeline stage
constexpr int kWgWin = kWgKC + 8;   // window rows: KC + 2 displaced rows, rounded to the 8-row swizzle group

struct WgradParams {
    int P, cin, cout, taps, Wp, n_kw, n_groups;
    int tiles_g, tiles_x, splits, steps_per_split, steps_total, n_work;
    int Cg;
    float* dw;
    int ld_dw;
    int atomic;
};

template <int BNX, int PASSES>
struct WgCfg {
    static constexpr int kOperands = (PASSES == 3) ? 2 : 1;
    static constexpr int kXRowBytes = BNX >= 64 ? 128 : BNX * 2;       // swizzle span of one X row chunk
    static constexpr int kXAtoms = BNX >= 64 ? BNX / 64 : 1;
    static constexpr int kXAtomBytes = kWgWin * kXRowBytes;             // one window box
    static constexpr int kXPlane = (kXAtoms * kXAtomBytes + 1023) / 1024 * 1024;
    static constexpr int kGAtomBytes = kWgKC * 128;                     // 64 pixels x 64 channels
    static constexpr int kGPlane = 2 * kGAtomBytes;
    static constexpr int kStageBytes = kOperands * (kGPlane + kXPlane);
    static constexpr int kBudget = 200 * 1024;
    static constexpr int kStagesRaw = kBudget / kStageBytes;
    static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
    static constexpr int kTapCols = BNX < 32 ? 32 : BNX;                // TMEM columns of one tap's accumulator
    static_assert(kStages >= 2, "need at least a double buffer");
};

template <int BNX, int PASSES>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap map_g_hi, const __grid_constant__ CUtensorMap map_g_lo,
             const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
             const WgradParams prm) {
    using Cfg = WgCfg<BNX, PASSES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + Cfg::kStages;
    uint64_t* tmem_full = empty_bar + Cfg::kStages;  // [2]
    uint64_t* tmem_empty = tmem_full + 2;            // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int acc_cols = prm.n_kw * Cfg::kTapCols;          // TMEM columns of one work item
    const int num_acc = (512 / acc_cols) >= 2 ? 2 : 1;      // double-buffer when it fits

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_g_hi);
        prefetch_tensormap(&map_x_hi);
        if (PASSES == 3) {
            prefetch_tensormap(&map_g_lo);
            prefetch_tensormap(&map_x_lo);
        }
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work item decode: x tile fastest, then g tile, then tap group, then row split
    auto decode = [&](int w, int& x0, int& g0, int& grp, int& step_begin, int& step_end) {
        x0 = (w % prm.tiles_x) * BNX;
        int r = w / prm.tiles_x;
        g0 = (r % prm.tiles_g) * 128;
        r /= prm.tiles_g;
        grp = r % prm.n_groups;
        const int sp = r / prm.n_groups;
        step_begin = sp * prm.steps_per_split;
        step_end = min(step_begin + prm.steps_per_split, prm.steps_total);
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int it = 0;
            for (int w = blockIdx.x; w < prm.n_work; w += gridDim.x) {
                int x0, g0, grp, sb, se;
                decode(w, x0, g0, grp, sb, se);
                // first tap of the group and its row shift (window row 0 = pixel p + shift)
                const int tap0 = grp * prm.n_kw;
                int shift = 0;
                if (prm.taps == 9) shift = (tap0 / 3 - 1) * prm.Wp + (tap0 % 3 - 1);
                const int g_atoms = (prm.Cg - g0) >= 128 ? 2 : 1;
                const uint32_t tx = Cfg::kOperands * (g_atoms * Cfg::kGAtomBytes + Cfg::kXAtoms * Cfg::kXAtomBytes);
                for (int st = sb; st < se; ++st, ++it) {
                    const int s = it % Cfg::kStages;
                    mbar_wait(&empty_bar[s], ((it / Cfg::kStages) & 1) ^ 1);
                    uint8_t* base = smem + s * Cfg::kStageBytes;
                    const int p0 = st * kWgKC;
                    mbar_arrive_expect_tx(&full_bar[s], tx);
                    for (int a = 0; a < g_atoms; ++a) {
                        tma_load_2d(base + a * Cfg::kGAtomBytes, &map_g_hi, &full_bar[s], g0 + a * 64, p0);
                        if (PASSES == 3)
                            tma_load_2d(base + Cfg::kGPlane + Cfg::kXPlane + a * Cfg::kGAtomBytes, &map_g_lo,
                                        &full_bar[s], g0 + a * 64, p0);
                    }
                    for (int a = 0; a < Cfg::kXAtoms; ++a) {
                        tma_load_2d(base + Cfg::kGPlane + a * Cfg::kXAtomBytes, &map_x_hi, &full_bar[s],
                                    x0 + a * 64, p0 + shift);
                        if (PASSES == 3)
                            tma_load_2d(base + 2 * Cfg::kGPlane + Cfg::kXPlane + a * Cfg::kXAtomBytes, &map_x_lo,
                                        &full_bar[s], x0 + a * 64, p0 + shift);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc_bf16_mn(128, BNX);
        constexpr uint32_t x_layout = BNX >= 64 ? 2u : (BNX == 32 ? 4u : 6u);   // SWIZZLE_128B / 64B / 32B (= the TMA box width)
        constexpr uint32_t x_sbo = 8 * Cfg::kXRowBytes;                     // 8 pixel rows
        constexpr uint32_t x_k16 = 16 * Cfg::kXRowBytes;                    // 16 pixels (one MMA K step)
        int it = 0, tl = 0;
        for (int w = blockIdx.x; w < prm.n_work; w += gridDim.x, ++tl) {
            int x0, g0, grp, sb, se;
            decode(w, x0, g0, grp, sb, se);
            const int acc = tl % num_acc;
            const uint32_t d_tmem = tmem_base + acc * acc_cols;
            mbar_wait(&tmem_empty[acc], ((tl / num_acc) & 1) ^ 1);
            tc_fence_after();
            for (int st = sb; st < se; ++st, ++it) {
                const int s = it % Cfg::kStages;
                mbar_wait(&full_bar[s], (it / Cfg::kStages) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t g_hi = smem_u32(smem + s * Cfg::kStageBytes);
                    const uint32_t x_hi = g_hi + Cfg::kGPlane;
                    const uint32_t g_lo = x_hi + Cfg::kXPlane;
                    const uint32_t x_lo = g_lo + Cfg::kGPlane;
                    for (int j = 0; j < prm.n_kw; ++j) {
                        const uint32_t dj = d_tmem + j * Cfg::kTapCols;
                        const uint32_t xoff = j * Cfg::kXRowBytes;  // displaced by j pixels
#pragma unroll
                        for (int k = 0; k < kWgKC / 16; ++k) {
                            const uint64_t da = make_mnmajor_desc(g_hi + k * 2048, Cfg::kGAtomBytes, 1024, 2u);
                            const uint64_t db = make_mnmajor_desc(x_hi + xoff + k * x_k16, Cfg::kXAtomBytes, x_sbo, x_layout);
                            mma_bf16_ss(dj, da, db, idesc, (st > sb || k > 0) ? 1u : 0u);
                            if (PASSES == 3) {
                                const uint64_t dal = make_mnmajor_desc(g_lo + k * 2048, Cfg::kGAtomBytes, 1024, 2u);
                                const uint64_t dbl = make_mnmajor_desc(x_lo + xoff + k * x_k16, Cfg::kXAtomBytes, x_sbo, x_layout);
                                mma_bf16_ss(dj, dal, db, idesc, 1u);
                                mma_bf16_ss(dj, da, dbl, idesc, 1u);
                            }
                        }
                    }
                    mma_commit(&empty_bar[s]);
                    if (st == se - 1) mma_commit(&tmem_full[acc]);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (warps 2..5): TMEM -> red.global.add / st.global =====================
        const int q = warp & 3;
        int tl = 0;
        for (int w = blockIdx.x; w < prm.n_work; w += gridDim.x, ++tl) {
            int x0, g0, grp, sb, se;
            decode(w, x0, g0, grp, sb, se);
            const int acc = tl % num_acc;
            const int n = g0 + q * 32 + lane;  // output channel of this thread's TMEM lane
            mbar_wait(&tmem_full[acc], (tl / num_acc) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * acc_cols + (uint32_t(q * 32) << 16);
            const int chunks = prm.n_kw * (Cfg::kTapCols / 32);
            for (int ch = 0; ch < chunks; ++ch) {
                uint32_t v[32];
                __syncwarp();
                tmem_ld_32x32(taddr + ch * 32, v);
                tmem_ld_wait();
                if (ch == chunks - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                }
                if (n >= prm.cout) continue;
                const int j = (ch * 32) / Cfg::kTapCols;
                const int c_base = x0 + (ch * 32) % Cfg::kTapCols;
                const int tap = grp * prm.n_kw + j;
                float* o = prm.dw + ((size_t)tap * prm.cin + c_base) * prm.ld_dw + n;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (i < BNX && c_base + i < prm.cin) {
                        const float val = __uint_as_float(v[i]);
                        if (prm.atomic) red_add_f32(o + (size_t)i * prm.ld_dw, val);
                        else o[(size_t)i * prm.ld_dw] = val;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int BNX, int PASSES>
static int launch_wgrad(const mv3d_wgrad_desc* d, cudaStream_t stream) {
    using Cfg = WgCfg<BNX, PASSES>;
    CUtensorMap mg_hi, mg_lo, mx_hi, mx_lo;
    int rc;
    const uint32_t xbox = BNX >= 64 ? 64 : BNX;
    if ((rc = make_map_2d(&mg_hi, d->d_g_hi, d->P, d->Cg, kWgKC, 64)) != MV3D_OK) return rc;
    if ((rc = make_map_2d(&mx_hi, d->d_x_hi, d->P, d->Cx, kWgWin, xbox)) != MV3D_OK) return rc;
    if (PASSES == 3) {
        if ((rc = make_map_2d(&mg_lo, d->d_g_lo, d->P, d->Cg, kWgKC, 64)) != MV3D_OK) return rc;
        if ((rc = make_map_2d(&mx_lo, d->d_x_lo, d->P, d->Cx, kWgWin, xbox)) != MV3D_OK) return rc;
    } else {
        mg_lo = mg_hi;
        mx_lo = mx_hi;
    }
    WgradParams p;
    p.P = d->P; p.cin = d->cin; p.cout = d->cout; p.taps = d->taps; p.Wp = d->Wp; p.Cg = d->Cg;
    p.n_kw = (d->taps == 9 && d->tap_window) ? 3 : 1;
    p.n_groups = d->taps / p.n_kw;
    p.tiles_g = ceil_div(d->cout, 128);
    p.tiles_x = ceil_div(d->Cx, BNX);
    p.steps_total = ceil_div(d->P, kWgKC);
    const int base_items = p.n_groups * p.tiles_g * p.tiles_x;
    int splits = d->split_rows > 0 ? d->split_rows : ceil_div(3 * num_sms(), base_items);
    const int max_splits = p.steps_total / 4 > 0 ? p.steps_total / 4 : 1;  // keep >= 4 pipeline steps per item
    if (splits > max_splits) splits = max_splits;
    if (!d->accumulate) splits = 1;
    if (splits < 1) splits = 1;
    p.steps_per_split = ceil_div(p.steps_total, splits);
    p.splits = ceil_div(p.steps_total, p.steps_per_split);
    p.atomic = d->accumulate ? 1 : 0;
    p.dw = d->d_dw;
    p.ld_dw = d->ld_dw > 0 ? d->ld_dw : d->cout;
    const long long n_work = (long long)base_items * p.splits;
    if (n_work > 0x7fffffffLL) return MV3D_ERR_ARG;
    p.n_work = (int)n_work;
    auto kern = wgrad_kernel<BNX, PASSES>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) { set_last_cuda_error(e); return MV3D_ERR_LAUNCH; }
        attr_set = true;
    }
    const int grid = p.n_work < num_sms() ? p.n_work : num_sms();
    kern<<<grid, kWgThreads, Cfg::kSmemBytes, stream>>>(mg_hi, mg_lo, mx_hi, mx_lo, p);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

template <int PASSES>
static int dispatch_wgrad(const mv3d_wgrad_desc* d, cudaStream_t s) {
    if (d->Cx % 128 == 0) return launch_wgrad<128, PASSES>(d, s);
    if (d->Cx % 64 == 0) return launch_wgrad<64, PASSES>(d, s);
    if (d->Cx == 32) return launch_wgrad<32, PASSES>(d, s);   // the im2col'd first image layer: taps = 1, K = 27 of 32
    return launch_wgrad<16, PASSES>(d, s);
}

}  // namespace mv3d

extern "C" __attribute__((visibility("default"))) int mv3d_conv_wgrad(const mv3d_wgrad_desc* d, void* stream) {
    using namespace mv3d;
    MV3D_REQUIRE(d != nullptr && d->P > 0 && d->cin > 0 && d->cout > 0);
    MV3D_REQUIRE(d->taps == 1 || d->taps == 9);
    MV3D_REQUIRE(d->taps == 1 || d->Wp > 1);
    MV3D_REQUIRE(d->passes == 1 || d->passes == 3);
    MV3D_REQUIRE(d->d_x_hi && d->d_g_hi && d->d_dw);
    MV3D_REQUIRE(d->passes == 1 || (d->d_x_lo && d->d_g_lo));
    MV3D_REQUIRE(d->Cg % 64 == 0 && d->Cg >= d->cout);
    MV3D_REQUIRE((d->Cx % 64 == 0 || d->Cx == 16 || (d->Cx == 32 && d->taps == 1)) && d->Cx >= d->cin);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return d->passes == 3 ? dispatch_wgrad<3>(d, s) : dispatch_wgrad<1>(d, s);
}
