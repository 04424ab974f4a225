// conv3x3 / 1x1 / fc as ONE implicit GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulator in
// TMEM, operands staged by TMA, mbarrier pipeline).  Replaces Network.conv / Network.fc of the
// reference (lib/networks/network.py:108-132,369-397 -> tf.nn.conv2d / xw_plus_b).
//
// Formulation.  Activations are stored zero-haloed and flattened ("PAD" layout, include/mv3d_b200.h):
// pixel index p = (b*Hp + h)*Wp + (w+1).  A 3x3 SAME stride-1 conv is then
//     D[p, n] = sum_{t=0..8} sum_c A[p + (t/3-1)*Wp + (t%3-1), c] * W[n, t*Cin + c]
// i.e. nine row-shifted GEMMs accumulated into the same TMEM tile.  Each k-step is one TMA box
// (128 pixel rows x KC channels, hardware 128B/32B swizzle) at a shifted row coordinate; rows before
// the first / after the last pixel are zero-filled by TMA, halo rows/columns hold zeros.  Output rows
// that are halo pixels are computed (<= 2/W + 1/H waste) and written as zeros by the epilogue.
//
// Tile: 128 (pixels) x BN (channels) per CTA, K step KC (64 -> SWIZZLE_128B, 16 -> SWIZZLE_32B).
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one elected lane),
//             warps 2..9 = epilogue (TMEM lane quarter = warp_id % 4, two warps per quarter on alternate column chunks).
// PASSES=3: x = hi + lo (bf16 pair); D += A_hi*W_hi + A_lo*W_hi + A_hi*W_lo  (fp32 accumulate).
// PASSES=2 (tap-reuse kernels only): "f16e5" operands (common.cuh): D += A_h*W_h (fp16) + [A_h8|A_l8]*[W_l8;W_h8] (e5m2,
//           kind::f8f6f4 at twice the rate) -- the same two correction terms for 2/3 of the tensor-pipe time.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace mv3d {

using namespace ptx;

constexpr int kBM = 128;
constexpr int kGemmThreads = 64 + 32 * 8;  // TMA warp, MMA warp, 8 epilogue warps

struct GemmParams {
    int M, N, Cin, taps, Hp, Wp, H, W;
    int k_chunks;       // Cin / KC
    int k_steps_total;  // taps * k_chunks
    int k_steps_per_split;
    int tiles_m, tiles_n, n_work;  // work item = (n tile fastest, m tile, k split)
    const float* bias;
    int relu;
    __nv_bfloat16* out_hi;
    __nv_bfloat16* out_lo;
    int ld_out;
    float* out_f32;
    int ld_f32;
    int f32_dense;
    int split_k;
    const __nv_bfloat16* mask_hi;  // backward-data epilogue: gate by (mask > 0), see mv3d_gemm_desc
    int ld_mask;
    float mask_scale;
    const float* addend;
    int ld_addend;
    int out_fmt;      // MV3D_FMT_BF16X2 / MV3D_FMT_F16E5 rendering of out_hi / out_lo
    float acc_scale;  // 2^-12 when the operands were f16e5 (PASSES == 2), else 1
    int dbg_flags;    // MV3D_GEMM_DBG (measurement only, results are garbage; pair kernel): 1 = the producer stops issuing TMA
                      // loads after the first lap of each ring (pure MMA rate on stale shared memory), 4 = the epilogue
                      // drains the accumulator but skips its math and stores.  Together they isolate the MMA rate.
    int softmax_cols;   // mv3d_gemm_desc::softmax_cols
    int pool, pool_Ho, pool_Wo, pool_nblk;   // fused 2x2 max-pool (pair kernel, POOL): pooled size, 128-column blocks per row
    int pdl;            // launched with programmaticStreamSerialization: griddepcontrol.wait before the first dependent read
    int k16_steps;      // 16-channel k-steps per 64-channel chunk that hold non-zero input channels (4 unless mv3d_gemm_desc::cin_valid)
    int e5_ksteps;      // f16e5 (PASSES = 2): 32-byte k-steps of the e5m2 row the MMA loop covers -- 4 = both correction terms, 2 = A_h W_l only
    long long* stamps;  // measurement only (mv3d_gemm_set_stamps): clock64 of pair 0's phases, see conv3x3_pair_kernel
};

__device__ __forceinline__ void stamp(const GemmParams& prm, int slot) {
    if (prm.stamps != nullptr && blockIdx.x == 0) prm.stamps[slot] = clock64();
}

template <int BN, int KC, int PASSES>
struct GemmCfg {
    static constexpr int kRowBytes = KC * 2;
    static constexpr int kABytes = kBM * kRowBytes;
    static constexpr int kBBytes = BN * kRowBytes;
    static constexpr int kOperands = (PASSES == 3) ? 2 : 1;
    static constexpr int kStageBytes = kOperands * (kABytes + kBBytes);
    static constexpr int kSmemBudget = 200 * 1024;
    static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + 4 * 512 /*bias*/;
    static constexpr int kAccCols = BN < 32 ? 32 : BN;   // TMEM columns of one accumulator
    static constexpr int kTmemCols = 2 * kAccCols;       // double-buffered: MMA of tile i+1 overlaps epilogue of tile i
    static_assert(kStages >= 2, "need at least a double buffer");
    static_assert(kTmemCols <= 512, "TMEM has 512 columns");
};

// Epilogue of one 128 x BN tile, executed by kEpiWarps = 8 epilogue warps: TMEM lane quarter q = warp % 4 (the
// hardware restriction), and the two warps of a quarter take alternate 32-column chunks (half = 0 / 1).
// TMEM -> registers -> scale / bias / ReLU -> operand rendering (bf16 hi/lo or f16e5) -> PAD layout and/or fp32.  Halo
// pixels are written as zeros; each warp hands the accumulator back as soon as its last chunk is in registers.
// (One warp per quarter and ~40 scalar instructions per element left conv4/conv5 -- a single tile per CTA pair -- with
// a 20 us serial tail and made the Cout = 64 layers epilogue-bound: ncu showed the MMAs at their floor and the tensor
// pipe idle the rest of the time.  Hence 8 warps, packed cvt instructions and vector bias loads.)
constexpr int kEpiWarps = 8;
constexpr int kBiasSmem = 512;   // bias vectors up to this many floats are staged in shared memory once per CTA

// All threads, before the set-up barrier: the epilogue then reads its bias from shared memory instead of paying an L2
// round trip per chunk on its serial TMEM-load -> render -> store chain.
__device__ __forceinline__ const float* stage_bias(const GemmParams& prm, float* bias_s) {
    if (prm.bias == nullptr || prm.N > kBiasSmem) return nullptr;
    for (int i = threadIdx.x; i < prm.N; i += blockDim.x) bias_s[i] = prm.bias[i];
    return bias_s;
}

__device__ __forceinline__ uint32_t cvt_pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t cvt_pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t cvt_pack_f16x2_sat(float lo, float hi) {   // |x| > 65504 -> +-65504 (= the clamp of split_f16e5)
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t cvt_e5m2x2_from_f16x2(uint32_t h2) {
    unsigned short r;
    asm("cvt.rn.satfinite.e5m2x2.f16x2 %0, %1;" : "=h"(r) : "r"(h2));
    return r;
}
__device__ __forceinline__ uint32_t cvt_pack_e5m2x2(float lo, float hi) {
    unsigned short r;
    asm("cvt.rn.satfinite.e5m2x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi), "f"(lo));
    return r;
}

// LEAN: the forward-inference instantiation -- no split-K, addend, gate or timing-experiment paths and only the vector
// stores (the host selects it when the descriptor allows); the dominant kernel's epilogue is a cold-code tail that
// stalls on instruction fetch, so its footprint matters.
template <int BN, int ACC_COLS, bool PAIR = false, bool LEAN = false>
__device__ __forceinline__ void epilogue_tile(const GemmParams& prm, uint32_t tmem_base, int acc, int m0, int n0, int q,
                                              int half, int lane, uint64_t* tmem_full, uint64_t* tmem_empty, int tl,
                                              const float* bias_s) {
    constexpr int kChunks = (BN + 31) / 32;
    const int row = q * 32 + lane;
    const long long p = (long long)m0 + row;
    const bool in_range = p < prm.M;
    bool halo = false;
    long long dense_row = p;
    if (prm.Hp > 0) {   // M < 2^31: 32-bit index arithmetic
        const unsigned pu = (unsigned)p;
        const unsigned t = pu / (unsigned)prm.Wp;
        const int wp = (int)(pu - t * (unsigned)prm.Wp);
        const unsigned b = t / (unsigned)prm.Hp;
        const int hp = (int)(t - b * (unsigned)prm.Hp);
        halo = (wp == 0) || (hp == prm.Hp - 1);
        dense_row = ((long long)b * prm.H + hp) * prm.W + (wp - 1);
    }
    const uint32_t taddr_row = tmem_base + acc * ACC_COLS + (uint32_t(q * 32) << 16);
    const int n_mine = (kChunks - half + 1) / 2;   // this warp's chunks: half, half + 2, ...
    // Hand the accumulator back to the MMA warp.  The tcgen05.ld results are in registers (wait::ld) and fenced, so a
    // RELAXED arrive is enough: a release at cluster scope compiles to MEMBAR.ALL.GPU, which stalls on this warp's own
    // global stores of the previous chunks / tile (~1 us per tile in the store-heavy early layers).
    auto release = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (PAIR) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tmem_empty[acc]), 0));  // the leader CTA's barrier
            else mbar_arrive(&tmem_empty[acc]);
        }
    };
    mbar_wait(&tmem_full[acc], (tl >> 1) & 1);
    tc_fence_after();
    if (n_mine == 0) {   // (BN = 32: one chunk) still one arrival per tile, in step with the accumulator phases
        release();
        return;
    }
    // One 32-column chunk: scale / bias / ReLU (/ addend / gate) -> operand rendering -> stores.
    auto process = [&](const uint32_t (&v)[32], const int c) {
    if (!in_range) return;
    if (prm.dbg_flags & 4) return;   // MV3D_GEMM_DBG=4 (timing experiment): drain the accumulator, skip the math / stores
    const int col0 = n0 + c;
    if (col0 >= prm.N) return;
    if (!LEAN && prm.split_k > 1) {
        if (halo) return;
        float* o = prm.out_f32 + (prm.f32_dense ? dense_row : p) * prm.ld_f32 + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (col0 + j < prm.N) atomicAdd(o + j, __uint_as_float(v[j]));
        return;
    }
    const bool full = LEAN || (col0 + 32 <= prm.N);
    float f[32];
    if (halo) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = 0.f;
    } else {
        const float sc = prm.acc_scale;
        const float* bsrc = bias_s != nullptr ? bias_s : prm.bias;   // generic loads: shared copy (N <= kBiasSmem) or global
        if (LEAN ? (prm.bias != nullptr) : (prm.bias != nullptr && full)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bsrc + col0 + j);
                // acc_scale is 1 or 2^-12: the product is exact, so the fused multiply-add rounds exactly like mul + add
                f[j] = __fmaf_rn(__uint_as_float(v[j]), sc, b4.x);
                f[j + 1] = __fmaf_rn(__uint_as_float(v[j + 1]), sc, b4.y);
                f[j + 2] = __fmaf_rn(__uint_as_float(v[j + 2]), sc, b4.z);
                f[j + 3] = __fmaf_rn(__uint_as_float(v[j + 3]), sc, b4.w);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x = __uint_as_float(v[j]) * sc;
                if (prm.bias != nullptr && col0 + j < prm.N) x += bsrc[col0 + j];
                f[j] = x;
            }
        }
        if (prm.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (!LEAN && col0 < prm.softmax_cols) {   // (bg, fg) score pairs -> probabilities; same arithmetic as softmax_pairs_kernel
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                if (col0 + j + 1 < prm.softmax_cols) {
                    const float m = fmaxf(f[j], f[j + 1]);
                    const float e0 = expf(f[j] - m), e1 = expf(f[j + 1] - m);
                    const float sm = e0 + e1;
                    f[j] = e0 / sm;
                    f[j + 1] = e1 / sm;
                }
            }
        }
    }
    if (!LEAN && prm.addend != nullptr && !halo) {  // second gradient path (dense rows), summed before the mask
        const float* ad = prm.addend + dense_row * prm.ld_addend + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (col0 + j < prm.N) f[j] += __ldg(ad + j);
    }
    if (!LEAN && prm.mask_hi != nullptr && !halo) {  // ReLU / dropout gate of the forward activation
        const __nv_bfloat16* mk = prm.mask_hi + p * prm.ld_mask + col0;
        if (full && (prm.ld_mask % 8 == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                const uint4 m4 = __ldg(reinterpret_cast<const uint4*>(mk + j));
                const uint32_t mw[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    // bf16 > 0  <=>  sign bit clear and magnitude bits non-zero
                    const uint32_t a = mw[e] & 0xFFFFu, b = mw[e] >> 16;
                    f[j + 2 * e] = (a != 0 && a < 0x8000u) ? f[j + 2 * e] * prm.mask_scale : 0.f;
                    f[j + 2 * e + 1] = (b != 0 && b < 0x8000u) ? f[j + 2 * e + 1] * prm.mask_scale : 0.f;
                }
            }
        } else {
            for (int j = 0; j < 32; ++j)
                if (col0 + j < prm.N) f[j] = (__bfloat162float(mk[j]) > 0.f) ? f[j] * prm.mask_scale : 0.f;
        }
    }
    if (prm.out_hi != nullptr && prm.out_fmt == MV3D_FMT_F16E5) {
        // fp16 plane + byte plane [e5m2(h) x64 | e5m2(residual * 4096) x64] per 64-channel chunk (host checks
        // N % 64 == 0, ld_out % 64 == 0, so every 32-column chunk is full and 16-byte aligned).  Same arithmetic as
        // split_f16e5 (common.cuh), two elements per cvt instruction.
        unsigned short* oh = reinterpret_cast<unsigned short*>(prm.out_hi) + p * prm.ld_out + col0;
        uint8_t* ob = reinterpret_cast<uint8_t*>(prm.out_lo) + p * prm.ld_out * 2 + f16e5_off(col0);
        uint32_t ph[16], p8[8], q8[8];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            uint32_t a8[2], b8[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                // (values beyond +-65504 saturate in the conversions: h = +-65504, residual +-57344 instead of the
                // clamped input's 0 -- only ever different for activations that overflow fp16)
                const float x0 = f[j + 2 * e], x1 = f[j + 2 * e + 1];
                const uint32_t h2 = cvt_pack_f16x2_sat(x0, x1);
                const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h2));
                ph[j / 2 + e] = h2;
                a8[e] = cvt_e5m2x2_from_f16x2(h2);
                b8[e] = cvt_pack_e5m2x2((x0 - hf.x) * kF16E5Scale, (x1 - hf.y) * kF16E5Scale);
            }
            p8[j / 4] = a8[0] | (a8[1] << 16);
            q8[j / 4] = b8[0] | (b8[1] << 16);
        }
        // each thread owns one pixel row: 256-bit stores = whole 32-byte sectors (rows are >= 128 B apart, so a
        // 16-byte store per lane would touch 32 half sectors per instruction -- the LSU transaction count, not the
        // arithmetic, was what the epilogue spent its time on)
        if (prm.dbg_flags & 16) {   // MV3D_GEMM_DBG=16 (timing experiment): render, keep the values live, store nothing
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) x ^= ph[j];
#pragma unroll
            for (int j = 0; j < 8; ++j) x ^= p8[j] ^ q8[j];
            if (x == 0x12345678u) oh[0] = 1;
            return;
        }
        st_global_v8(oh, ph);
        st_global_v8(oh + 16, ph + 8);
        st_global_v8(ob, p8);
        st_global_v8(ob + 64, q8);
    } else if (prm.out_hi != nullptr) {
        __nv_bfloat16* oh = prm.out_hi + p * prm.ld_out + col0;
        __nv_bfloat16* ol = prm.out_lo ? prm.out_lo + p * prm.ld_out + col0 : nullptr;
        if (LEAN || (full && !(prm.dbg_flags & 8) && (prm.ld_out % 16 == 0) && ((reinterpret_cast<uintptr_t>(prm.out_hi) | reinterpret_cast<uintptr_t>(prm.out_lo)) & 31) == 0)) {
            uint32_t ph[16], pl[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {   // = split_bf16 on two elements
                const float x0 = f[2 * e], x1 = f[2 * e + 1];
                const uint32_t h2 = cvt_pack_bf16x2(x0, x1);
                ph[e] = h2;
                pl[e] = cvt_pack_bf16x2(x0 - __uint_as_float(h2 << 16), x1 - __uint_as_float(h2 & 0xFFFF0000u));
            }
            st_global_v8(oh, ph);
            st_global_v8(oh + 16, ph + 8);
            if (ol) {
                st_global_v8(ol, pl);
                st_global_v8(ol + 16, pl + 8);
            }
        } else if (full && (prm.ld_out % 8 == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint32_t ph4[4], pl4[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float x0 = f[j + 2 * e], x1 = f[j + 2 * e + 1];
                    const uint32_t h2 = cvt_pack_bf16x2(x0, x1);
                    ph4[e] = h2;
                    pl4[e] = cvt_pack_bf16x2(x0 - __uint_as_float(h2 << 16), x1 - __uint_as_float(h2 & 0xFFFF0000u));
                }
                *reinterpret_cast<uint4*>(oh + j) = make_uint4(ph4[0], ph4[1], ph4[2], ph4[3]);
                if (ol) *reinterpret_cast<uint4*>(ol + j) = make_uint4(pl4[0], pl4[1], pl4[2], pl4[3]);
            }
        } else {
            for (int j = 0; j < 32 && col0 + j < prm.N; ++j) {
                __nv_bfloat16 h, l;
                split_bf16(f[j], h, l);
                oh[j] = h;
                if (ol) ol[j] = l;
            }
        }
    }
    if (prm.out_f32 != nullptr && !(halo && prm.f32_dense)) {
        float* o = prm.out_f32 + (prm.f32_dense ? dense_row : p) * prm.ld_f32 + col0;
        if (LEAN || (full && !(prm.dbg_flags & 8) && (prm.ld_f32 % 8 == 0) && (reinterpret_cast<uintptr_t>(prm.out_f32) & 31) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) st_global_v8(o + j, reinterpret_cast<const uint32_t*>(f + j));
        } else if (full && (prm.ld_f32 % 4 == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
            for (int j = 0; j < 32 && col0 + j < prm.N; ++j) o[j] = f[j];
        }
    }
    };
    // Two register buffers: the TMEM load of chunk k+1 is in flight while chunk k is rendered and stored (one tile per
    // CTA pair in the 512-channel layers: the epilogue is an exposed, latency-bound tail).
    uint32_t va[32], vb[32];
    __syncwarp();
    tmem_ld_32x32(taddr_row + half * 32, va);
    tmem_ld_wait_dep(va);
#pragma unroll 1
    for (int k = 0; k < n_mine; ++k) {
        __syncwarp();
        if (k + 1 < n_mine) tmem_ld_32x32(taddr_row + (half + 2 * (k + 1)) * 32, vb);
        else release();
        process(va, (half + 2 * k) * 32);          // ONE inlined copy of the rendering code (instruction-cache footprint)
        if (k + 1 < n_mine) {
            tmem_ld_wait_dep(vb);
#pragma unroll
            for (int j = 0; j < 32; ++j) va[j] = vb[j];
        }
    }
}

// Work item -> first output pixel (flat PAD index) of CTA `rank`.  Plain: 256 consecutive pixels per pair.  POOL: item =
// (frame b, row pair i, 128-column block k); CTA r owns pixels (2i + r, 128k .. 128k + 127).
__device__ __forceinline__ int tile_m0(const GemmParams& prm, int w, int tiles_n, int rank, int& pb, int& pi, int& pk) {
    if (prm.pool) {
        pk = w % prm.pool_nblk;
        const int t = w / prm.pool_nblk;
        pi = t % prm.pool_Ho;
        pb = t / prm.pool_Ho;
        return (pb * prm.Hp + 2 * pi + rank) * prm.Wp + 128 * pk + 1;
    }
    pb = pi = pk = 0;
    return (w / tiles_n) * (2 * kBM) + rank * kBM;
}

// Epilogue of one POOL tile (see PairCfg).  Warp (q, half) reads its TMEM quarter's chunks half, half + 2, ...:
// bias + ReLU, horizontal max of the pixel pair (lanes 2j, 2j+1).  The CTA of rank r FINALISES the chunks whose warps
// have half == r: the other CTA's warps of the same (q, half) -- holding the other image row -- send their 16 x 32
// values per chunk into this CTA's exchange buffer (st.async, bytes counted on x_full); the finaliser takes the vertical
// max, renders the operand format and stores the pooled pixel (+ the zero halo column / row of the pooled PAD tensor).
// x_empty (in the sender's CTA, one arrival per finalising warp of the peer) frees the buffer two tiles later.
template <int BN, int ACC_COLS>
__device__ __forceinline__ void epilogue_tile_pool(const GemmParams& prm, uint32_t tmem_base, int acc, int pb, int pi,
                                                   int pk, int rank, int q, int half, int lane, uint64_t* tmem_full,
                                                   uint64_t* tmem_empty, int tl, const float* bias_s, float* xbuf,
                                                   uint64_t* x_full, uint64_t* x_empty) {
    constexpr int kMine = BN / 64;                      // chunks per warp
    const bool finalizer = (half == rank);
    const int ph = tl & 1;
    const uint32_t par = (tl >> 1) & 1;
    const uint32_t peer = (uint32_t)(rank ^ 1);
    if (finalizer && q == 0 && lane == 0) mbar_arrive_expect_tx(&x_full[ph], kMine * 8192);
    mbar_wait(&tmem_full[acc], par);
    tc_fence_after();
    if (!finalizer) mbar_wait(&x_empty[ph], par ^ 1);   // the peer has read what was sent two tiles ago
    const uint32_t taddr_row = tmem_base + acc * ACC_COLS + (uint32_t(q * 32) << 16);
    const float sc = prm.acc_scale;
    const float* bsrc = bias_s != nullptr ? bias_s : prm.bias;
    const int Ho = prm.pool_Ho, Wo = prm.pool_Wo;
    const int jj = 64 * pk + 16 * q + (lane >> 1);      // pooled column of this lane pair
    const bool even = (lane & 1) == 0;
#pragma unroll 1
    for (int k = 0; k < kMine; ++k) {
        const int c = (half + 2 * k) * 32;              // first channel of the chunk (n0 = 0: N == BN)
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(taddr_row + c, v);
        tmem_ld_wait();
        if (k == kMine - 1) {                           // accumulator drained by this warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
        }
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(v[j]);
            x = (prm.bias != nullptr) ? __fmaf_rn(x, sc, bsrc[c + j]) : x * sc;
            if (prm.relu) x = fmaxf(x, 0.f);
            f[j] = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 1));        // max over the pixel pair (w, w+1), w even
        }
        float* slot = xbuf + ((size_t)((ph * kMine + k) * 4 + q) * 16 + (lane >> 1)) * 32;
        if (!finalizer) {
            if (even) {
                const uint32_t dst = mapa_u32(smem_u32(slot), peer), bar = mapa_u32(smem_u32(&x_full[ph]), peer);
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    st_async_v4(dst + j * 4, __float_as_uint(f[j]), __float_as_uint(f[j + 1]), __float_as_uint(f[j + 2]),
                                __float_as_uint(f[j + 3]), bar);
            }
            continue;
        }
        if (k == 0) mbar_wait(&x_full[ph], par);        // every chunk of this tile has landed
        if (!even || jj >= Wo) continue;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 o = *reinterpret_cast<const float4*>(slot + j);
            f[j] = fmaxf(f[j], o.x); f[j + 1] = fmaxf(f[j + 1], o.y);
            f[j + 2] = fmaxf(f[j + 2], o.z); f[j + 3] = fmaxf(f[j + 3], o.w);
        }
        // pooled pixel (pi, jj) of frame pb in the pooled PAD tensor, and the halo positions this thread zeroes
        const long long row0 = ((long long)pb * (Ho + 1) + pi) * (Wo + 1);
        const long long qo = row0 + jj + 1;
        uint32_t zero8[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        if (prm.out_fmt == MV3D_FMT_F16E5) {
            unsigned short* oh = reinterpret_cast<unsigned short*>(prm.out_hi);
            uint8_t* ob = reinterpret_cast<uint8_t*>(prm.out_lo);
            uint32_t ph16[16], p8[8], q8[8];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                uint32_t a8[2], b8[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float x0 = fminf(f[j + 2 * e], 65504.f), x1 = fminf(f[j + 2 * e + 1], 65504.f);
                    const float y0 = fmaxf(x0, -65504.f), y1 = fmaxf(x1, -65504.f);
                    const uint32_t h2 = cvt_pack_f16x2(y0, y1);
                    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h2));
                    ph16[j / 2 + e] = h2;
                    a8[e] = cvt_pack_e5m2x2(hf.x, hf.y);
                    b8[e] = cvt_pack_e5m2x2((y0 - hf.x) * kF16E5Scale, (y1 - hf.y) * kF16E5Scale);
                }
                p8[j / 4] = a8[0] | (a8[1] << 16);
                q8[j / 4] = b8[0] | (b8[1] << 16);
            }
            auto put = [&](long long pix, const uint32_t* h, const uint32_t* a, const uint32_t* b) {
                unsigned short* dh = oh + pix * prm.ld_out + c;
                uint8_t* db = ob + pix * prm.ld_out * 2 + f16e5_off(c);
                st_global_v8(dh, h);
                st_global_v8(dh + 16, h + 8);
                st_global_v8(db, a);
                st_global_v8(db + 64, b);
            };
            uint32_t zero16[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) zero16[e] = 0u;
            put(qo, ph16, p8, q8);
            if (jj == 0) put(row0, zero16, zero8, zero8);
            if (pi == Ho - 1) {
                put(qo + (Wo + 1), zero16, zero8, zero8);
                if (jj == 0) put(row0 + (Wo + 1), zero16, zero8, zero8);
            }
        } else {
            __nv_bfloat16* oh = prm.out_hi;
            __nv_bfloat16* ol = prm.out_lo;
            uint32_t phb[16], plb[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const float x0 = f[2 * e], x1 = f[2 * e + 1];
                const uint32_t h2 = cvt_pack_bf16x2(x0, x1);
                phb[e] = h2;
                plb[e] = cvt_pack_bf16x2(x0 - __uint_as_float(h2 << 16), x1 - __uint_as_float(h2 & 0xFFFF0000u));
            }
            auto put = [&](long long pix, const uint32_t* h, const uint32_t* l) {
                st_global_v8(oh + pix * prm.ld_out + c, h);
                st_global_v8(oh + pix * prm.ld_out + c + 16, h + 8);
                if (ol) {
                    st_global_v8(ol + pix * prm.ld_out + c, l);
                    st_global_v8(ol + pix * prm.ld_out + c + 16, l + 8);
                }
            };
            uint32_t zero16[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) zero16[e] = 0u;
            put(qo, phb, plb);
            if (jj == 0) put(row0, zero16, zero16);
            if (pi == Ho - 1) {
                put(qo + (Wo + 1), zero16, zero16);
                if (jj == 0) put(row0 + (Wo + 1), zero16, zero16);
            }
        }
    }
    if (finalizer) {   // this warp has read its slots of the buffer: the peer may refill it
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_cta(mapa_u32(smem_u32(&x_empty[ph]), peer));
    }
}

// Persistent: grid = min(#work items, #SMs); CTA c processes work items c, c+grid, ... where a work item is
// (n tile fastest, m tile, k split).  Three asynchronous pipelines run concurrently inside a CTA:
//   TMA producer  --full/empty mbarriers (smem ring, continuous across tiles)-->  MMA issuer
//   MMA issuer    --tmem_full/tmem_empty mbarriers (2 accumulators)-->            epilogue warps
template <int BN, int KC, int PASSES>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                 const GemmParams prm) {
    using Cfg = GemmCfg<BN, KC, PASSES>;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operand tiles need 1024-byte alignment.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + Cfg::kStages;
    uint64_t* tmem_full = empty_bar + Cfg::kStages;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;             // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    const float* bias_s = stage_bias(prm, reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_a_hi);
        prefetch_tensormap(&map_w_hi);
        if (PASSES == 3) {
            prefetch_tensormap(&map_a_lo);
            prefetch_tensormap(&map_w_lo);
        }
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], kEpiWarps);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_n = prm.tiles_n, tiles_m = prm.tiles_m;
    const int n_work = prm.n_work;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int it = 0;  // smem ring position, continuous across work items
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int n0 = (w % tiles_n) * BN;
                const int r = w / tiles_n;
                const int m0 = (r % tiles_m) * kBM;
                const int k_begin = (r / tiles_m) * prm.k_steps_per_split;
                const int k_end = min(k_begin + prm.k_steps_per_split, prm.k_steps_total);
                for (int ks = k_begin; ks < k_end; ++ks, ++it) {
                    const int s = it % Cfg::kStages;
                    const uint32_t ph = (it / Cfg::kStages) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    const int tap = ks / prm.k_chunks;
                    const int c0 = (ks - tap * prm.k_chunks) * KC;
                    int shift = 0;
                    if (prm.taps == 9) shift = (tap / 3 - 1) * prm.Wp + (tap % 3 - 1);
                    uint8_t* st = stage_base + s * Cfg::kStageBytes;
                    mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
                    tma_load_2d(st, &map_a_hi, &full_bar[s], c0, m0 + shift);
                    tma_load_2d(st + Cfg::kABytes, &map_w_hi, &full_bar[s], tap * prm.Cin + c0, n0);
                    if (PASSES == 3) {
                        tma_load_2d(st + Cfg::kABytes + Cfg::kBBytes, &map_a_lo, &full_bar[s], c0, m0 + shift);
                        tma_load_2d(st + 2 * Cfg::kABytes + Cfg::kBBytes, &map_w_lo, &full_bar[s], tap * prm.Cin + c0,
                                    n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc_bf16(kBM, BN);
        int it = 0, tl = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++tl) {
            const int r = w / tiles_n;
            const int k_begin = (r / tiles_m) * prm.k_steps_per_split;
            const int k_end = min(k_begin + prm.k_steps_per_split, prm.k_steps_total);
            const int acc = tl & 1;
            const uint32_t d_tmem = tmem_base + acc * Cfg::kAccCols;
            mbar_wait(&tmem_empty[acc], ((tl >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
            tc_fence_after();
            for (int ks = k_begin; ks < k_end; ++ks, ++it) {
                const int s = it % Cfg::kStages;
                const uint32_t ph = (it / Cfg::kStages) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_hi = smem_u32(stage_base + s * Cfg::kStageBytes);
                    const uint32_t w_hi = a_hi + Cfg::kABytes;
                    const uint32_t a_lo = w_hi + Cfg::kBBytes;
                    const uint32_t w_lo = a_lo + Cfg::kABytes;
#pragma unroll
                    for (int k = 0; k < KC / 16; ++k) {
                        const uint32_t koff = k * 32;  // 16 bf16 along K inside the swizzled row
                        const uint64_t da = make_kmajor_desc(a_hi + koff, Cfg::kRowBytes);
                        const uint64_t db = make_kmajor_desc(w_hi + koff, Cfg::kRowBytes);
                        mma_bf16_ss(d_tmem, da, db, idesc, (ks > k_begin || k > 0) ? 1u : 0u);
                        if (PASSES == 3) {
                            const uint64_t dal = make_kmajor_desc(a_lo + koff, Cfg::kRowBytes);
                            const uint64_t dbl = make_kmajor_desc(w_lo + koff, Cfg::kRowBytes);
                            mma_bf16_ss(d_tmem, dal, db, idesc, 1u);
                            mma_bf16_ss(d_tmem, da, dbl, idesc, 1u);
                        }
                    }
                    mma_commit(&empty_bar[s]);                       // frees the smem slot when these MMAs retire
                    if (ks == k_end - 1) mma_commit(&tmem_full[acc]);  // accumulator complete
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int q = warp & 3;  // TMEM lane quarter this warp may read
        int tl = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++tl) {
            const int n0 = (w % tiles_n) * BN;
            const int m0 = ((w / tiles_n) % tiles_m) * kBM;
            epilogue_tile<BN, Cfg::kAccCols>(prm, tmem_base, tl & 1, m0, n0, q, (warp - 2) >> 2, lane, tmem_full, tmem_empty, tl, bias_s);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}


// ------------------------------------------------------------------------------------------------
// 3x3 conv with TAP REUSE.  The three kw taps of one kernel row read the same activation rows displaced by one
// pixel.  Instead of three 128-row TMA boxes the producer fetches ONE 136-row box (rows m0-1 .. m0+134 of the
// shifted image row) and the MMA issuer addresses it three times with the descriptor start advanced by kw rows
// (128 B each).  The tensor core applies the 128B-swizzle XOR to the absolute shared-memory address bits, exactly
// as TMA did when it wrote the box, so a start displaced by whole rows needs no descriptor base_offset (measured
// on B200: base_offset = kw gives wrong results, 0 is bit-identical to the three-box kernel).  Activation traffic from L2
// drops 3x -- the early, L2-bound layers (Cout = 64/128) are bounded by exactly that traffic.
// Two rings: A entries (one per (cin chunk, kh)) and W entries (one per (cin chunk, kh, kw)).
// ------------------------------------------------------------------------------------------------
constexpr int kReuseRows = 136;  // 128 + 2 shifted rows, rounded to the 8-row swizzle group

template <int BN, int PASSES>
struct ReuseCfg {
    static constexpr int kOperands = (PASSES >= 2) ? 2 : 1;
    static constexpr int kAPlane = 18 * 1024;                 // 136 rows x 128 B = 17408, padded to 1024 multiple
    static constexpr int kABoxBytes = kReuseRows * 128;
    static constexpr int kAEntry = kAPlane * kOperands;
    static constexpr int kWPlane = BN * 128;
    static constexpr int kWEntry = kWPlane * kOperands;
    static constexpr int kNA = (BN <= 64 ? 3 : 2) * (PASSES >= 2 ? 1 : 2);
    static constexpr int kBudget = 200 * 1024;
    static constexpr int kNWRaw = (kBudget - kNA * kAEntry) / kWEntry;
    static constexpr int kNW = kNWRaw > 8 ? 8 : kNWRaw;
    static constexpr int kSmemBytes = kNA * kAEntry + kNW * kWEntry + 1024 + 512 + 4 * 512 /*bias*/;
    static constexpr int kAccCols = BN < 32 ? 32 : BN;
    static constexpr int kTmemCols = 2 * kAccCols;
    static_assert(kNW >= 3, "W ring too shallow");
    static_assert(kTmemCols <= 512, "TMEM has 512 columns");
};

template <int BN, int PASSES>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv3x3_reuse_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                     const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                     const GemmParams prm) {
    using Cfg = ReuseCfg<BN, PASSES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_ring = smem;
    uint8_t* w_ring = smem + Cfg::kNA * Cfg::kAEntry;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(w_ring + Cfg::kNW * Cfg::kWEntry);
    uint64_t* a_empty = a_full + Cfg::kNA;
    uint64_t* w_full = a_empty + Cfg::kNA;
    uint64_t* w_empty = w_full + Cfg::kNW;
    uint64_t* tmem_full = w_empty + Cfg::kNW;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    const float* bias_s = stage_bias(prm, reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a_full) + 512));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_a_hi);
        prefetch_tensormap(&map_w_hi);
        if (PASSES >= 2) {
            prefetch_tensormap(&map_a_lo);
            prefetch_tensormap(&map_w_lo);
        }
        for (int i = 0; i < Cfg::kNA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < Cfg::kNW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], kEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_n = prm.tiles_n, tiles_m = prm.tiles_m, n_work = prm.n_work;
    const int n_groups = prm.k_chunks * 3;  // (cin chunk, kh) groups per tile, 3 kw taps each

    if (warp == 0) {
        if (lane == 0) {
            int ia = 0, iw = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int n0 = (w % tiles_n) * BN;
                const int m0 = ((w / tiles_n) % tiles_m) * kBM;
                for (int g = 0; g < n_groups; ++g, ++ia) {
                    const int chunk = g / 3, kh = g - chunk * 3;
                    const int c0 = chunk * 64;
                    const int ea = ia % Cfg::kNA;
                    mbar_wait(&a_empty[ea], ((ia / Cfg::kNA) & 1) ^ 1);
                    uint8_t* ab = a_ring + ea * Cfg::kAEntry;
                    mbar_arrive_expect_tx(&a_full[ea], Cfg::kABoxBytes * Cfg::kOperands);
                    const int arow = m0 + (kh - 1) * prm.Wp - 1;
                    tma_load_2d(ab, &map_a_hi, &a_full[ea], c0, arow);
                    if (PASSES >= 2) tma_load_2d(ab + Cfg::kAPlane, &map_a_lo, &a_full[ea], c0, arow);
                    for (int kw = 0; kw < 3; ++kw, ++iw) {
                        const int ew = iw % Cfg::kNW;
                        mbar_wait(&w_empty[ew], ((iw / Cfg::kNW) & 1) ^ 1);
                        uint8_t* wb = w_ring + ew * Cfg::kWEntry;
                        mbar_arrive_expect_tx(&w_full[ew], Cfg::kWEntry);
                        const int kcol = (kh * 3 + kw) * prm.Cin + c0;
                        tma_load_2d(wb, &map_w_hi, &w_full[ew], kcol, n0);
                        if (PASSES >= 2) tma_load_2d(wb + Cfg::kWPlane, &map_w_lo, &w_full[ew], kcol, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = PASSES == 2 ? make_idesc_f16(kBM, BN) : make_idesc_bf16(kBM, BN);
        constexpr uint32_t idesc8 = make_idesc_e5m2(kBM, BN);
        (void)idesc8;
        int ia = 0, iw = 0, tl = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++tl) {
            const int acc = tl & 1;
            const uint32_t d_tmem = tmem_base + acc * Cfg::kAccCols;
            mbar_wait(&tmem_empty[acc], ((tl >> 1) & 1) ^ 1);
            tc_fence_after();
            for (int g = 0; g < n_groups; ++g, ++ia) {
                const int ea = ia % Cfg::kNA;
                mbar_wait(&a_full[ea], (ia / Cfg::kNA) & 1);
                const uint32_t a_hi = smem_u32(a_ring + ea * Cfg::kAEntry);
                const uint32_t a_lo = a_hi + Cfg::kAPlane;
                for (int kw = 0; kw < 3; ++kw, ++iw) {
                    const int ew = iw % Cfg::kNW;
                    mbar_wait(&w_full[ew], (iw / Cfg::kNW) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t w_hi = smem_u32(w_ring + ew * Cfg::kWEntry);
                        const uint32_t w_lo = w_hi + Cfg::kWPlane;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t aoff = kw * 128 + k * 32;  // kw rows down, 16 bf16 along K
                            const uint64_t da = make_kmajor_desc(a_hi + aoff, 128);
                            const uint64_t db = make_kmajor_desc(w_hi + k * 32, 128);
                            mma_bf16_ss(d_tmem, da, db, idesc, (g > 0 || kw > 0 || k > 0) ? 1u : 0u);
                            if (PASSES == 3) {
                                const uint64_t dal = make_kmajor_desc(a_lo + aoff, 128);
                                const uint64_t dbl = make_kmajor_desc(w_lo + k * 32, 128);
                                mma_bf16_ss(d_tmem, dal, db, idesc, 1u);
                                mma_bf16_ss(d_tmem, da, dbl, idesc, 1u);
                            }
                            if (PASSES == 2) {  // 32 e5m2 of the byte plane: [A_h8|A_l8] . [W_l8;W_h8]
                                const uint64_t dal = make_kmajor_desc(a_lo + aoff, 128);
                                const uint64_t dbl = make_kmajor_desc(w_lo + k * 32, 128);
                                mma_f8_ss(d_tmem, dal, dbl, idesc8, 1u);
                            }
                        }
                        mma_commit(&w_empty[ew]);
                        if (kw == 2) mma_commit(&a_empty[ea]);
                        if (kw == 2 && g == n_groups - 1) mma_commit(&tmem_full[acc]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        const int q = warp & 3;
        int tl = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++tl) {
            const int n0 = (w % tiles_n) * BN;
            const int m0 = ((w / tiles_n) % tiles_m) * kBM;
            epilogue_tile<BN, Cfg::kAccCols>(prm, tmem_base, tl & 1, m0, n0, q, (warp - 2) >> 2, lane, tmem_full, tmem_empty, tl, bias_s);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
// 3x3 conv, tap reuse, CTA PAIR (tcgen05 cta_group::2).  Two CTAs of a cluster (one TPC) compute ONE 256 x BN tile:
// CTA r stages activation rows [m0 + 128 r, +128) (the same 136-row displaced box as above) and weight rows
// [n0 + r BN/2, +BN/2); the leader (rank 0) issues M=256 MMAs that read both CTAs' shared memory and write a
// 128 x BN accumulator into EACH CTA's TMEM.  Per output element each SM now pulls half the weight bytes from L2 and
// reads half the B operand from shared memory: the 128x128 single-CTA tile needs 57 B/clk/SM from L2 at full
// tensor rate (the chip delivers ~43), this one 28.
// Protocol: TMA of both CTAs completes on the LEADER's full barriers (expect_tx = both halves); tcgen05.commit
// multicasts the "slot free" / "accumulator ready" arrivals to both CTAs; the epilogue warps of both CTAs
// arrive on the leader's tmem_empty barrier (count 8).
// ------------------------------------------------------------------------------------------------
// WRES (64 -> 64 channel layers: conv1_2, BEV conv1_1): all nine weight taps of the layer (72 KB per CTA) are loaded
// ONCE and stay resident while the persistent CTA pair walks its ~30 tiles.  These layers are bound by the L2 -> SM
// operand feed (176 KB per tile and CTA at ~37 B/clk/SM of the ~43 the chip delivers); the weights were 41 % of it.
// POOL (conv1_2 / conv2_2 followed by Network.max_pool(2,2,2,2,'VALID')): the pair tile is 2 image rows x 128 columns
// (CTA r = row 2i + r), each CTA takes the horizontal pair maximum with a lane shuffle and the two CTAs exchange half
// of their channel chunks through distributed shared memory (st.async + complete_tx), so every 2x2 window is reduced
// on chip and only the pooled activation (1/4 of the bytes) is written -- these layers are bound by their HBM write.
template <int BN, int PASSES, bool WRES = false, bool POOL = false>
struct PairCfg {
    static constexpr int kOperands = (PASSES >= 2) ? 2 : 1;
    static constexpr int kAPlane = 18 * 1024;
    static constexpr int kABoxBytes = kReuseRows * 128;
    static constexpr int kAEntry = kAPlane * kOperands;
    static constexpr int kWRows = BN / 2;                      // weight rows staged by each CTA
    static constexpr int kWPlane = kWRows * 128;
    static constexpr int kWEntry = kWPlane * kOperands;
    static constexpr int kNA = (BN <= 64 ? 3 : 2) * (PASSES >= 2 ? 1 : 2);
    static constexpr int kBudget = 200 * 1024;
    static constexpr int kXChunks = BN / 64;                          // channel chunks a warp sends / receives per tile
    static constexpr int kXPhase = kXChunks * 8192;                   // exchange bytes per tile: 4 warps x 16 columns x 32 ch x 4 B
    static constexpr int kXBytes = POOL ? 2 * kXPhase : 0;            // double-buffered
    static constexpr int kNWRaw = (kBudget - kNA * kAEntry - kXBytes) / kWEntry;
    static constexpr int kNW = WRES ? 9 : (kNWRaw > 8 ? 8 : kNWRaw);   // WRES: entry = tap, never recycled
    static constexpr int kSmemBytes = kNA * kAEntry + kNW * kWEntry + kXBytes + 1024 + 512 + 4 * 512 /*bias*/;
    static constexpr int kAccCols = BN < 32 ? 32 : BN;
    static constexpr int kTmemCols = 2 * kAccCols;
    static_assert(kNW >= 3, "W ring too shallow");
    static_assert(kSmemBytes <= 227 * 1024, "shared memory");
    static_assert(kTmemCols <= 512, "TMEM has 512 columns");
    static_assert(BN % 16 == 0 && BN <= 256, "M=256 MMA: N multiple of 16, at most 256");
};

// Register budget of the pair kernel.  Default: whatever the epilogue wants (168 x 320 threads = 82 % of the register file:
// hardly anything of another stream can share the SM).  -DMV3D_PAIR_MAXNREG=128 caps it (a few spilled epilogue values)
// so that small kernels of the other trunk / frame (<= 24 k registers, <= 27 KB shared memory) co-reside.
#ifdef MV3D_PAIR_MAXNREG
#define MV3D_PAIR_BOUNDS __maxnreg__(MV3D_PAIR_MAXNREG)
#else
#define MV3D_PAIR_BOUNDS __launch_bounds__(kGemmThreads, 1)
#endif
template <int BN, int PASSES, bool LEAN = false, bool WRES = false, bool POOL = false>
__global__ void __cluster_dims__(2, 1, 1) MV3D_PAIR_BOUNDS
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                    const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                    const GemmParams prm) {
    using Cfg = PairCfg<BN, PASSES, WRES, POOL>;
    extern __shared__ uint8_t smem_raw[];
    // the dynamic window starts at the same offset in both CTAs, so the aligned pointers are at equal offsets too
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_ring = smem;
    uint8_t* w_ring = smem + Cfg::kNA * Cfg::kAEntry;
    float* xbuf = reinterpret_cast<float*>(w_ring + Cfg::kNW * Cfg::kWEntry);   // POOL: exchange buffer (else empty)
    uint64_t* a_full = reinterpret_cast<uint64_t*>(w_ring + Cfg::kNW * Cfg::kWEntry + Cfg::kXBytes);
    uint64_t* a_empty = a_full + Cfg::kNA;
    uint64_t* w_full = a_empty + Cfg::kNA;
    uint64_t* w_empty = w_full + Cfg::kNW;
    uint64_t* tmem_full = w_empty + Cfg::kNW;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* x_full = tmem_empty + 2;     // [2] POOL
    uint64_t* x_empty = x_full + 2;        // [2] POOL
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_empty + 2);
    const float* bias_s = stage_bias(prm, reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a_full) + 512));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) stamp(prm, 0);   // kernel start
    if (prm.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the NEXT kernel may be scheduled as SMs free up

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_a_hi);
        prefetch_tensormap(&map_w_hi);
        if (PASSES >= 2) {
            prefetch_tensormap(&map_a_lo);
            prefetch_tensormap(&map_w_lo);
        }
        for (int i = 0; i < Cfg::kNA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < Cfg::kNW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 2 * kEpiWarps); }
        for (int a = 0; a < 2; ++a) { mbar_init(&x_full[a], 1); mbar_init(&x_empty[a], 4); }   // POOL exchange
        fence_barrier_init();
    }
    // tcgen05.alloc.cta_group::2 is a two-party protocol that ptxas expands into messages through the RESERVED shared
    // memory of both CTAs (UTCATOMSWS.2CTA.FIND_AND_SET in one CTA, an mbarrier hand-off of the address to the other).
    // Both CTAs must therefore be running before either starts it: a cluster barrier first.  Without it the kernel
    // hangs every few thousand launches under multi-stream load (cuda-gdb: CTA 0 past its alloc, lane 0 of CTA 1's
    // warp 1 spinning inside the alloc sequence).  For the same reason the pair's allocation permit is only given up
    // after both allocs have returned, and no CTA exits before both have run the dealloc sequence.
    cluster_sync_relaxed();   // execution only: both CTAs are running
    if (warp == 1) tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    tc_fence_before();
    cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / TMA completion
    tc_fence_after();
    if (warp == 1) tmem_relinquish_pair();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_n = prm.tiles_n, n_work = prm.n_work;
    const int n_groups = prm.k_chunks * 3;
    const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    if (threadIdx.x == 0) stamp(prm, 1);   // set-up done (barriers, cluster syncs, TMEM)

    if (warp == 0) {
        if (lane == 0) {
            // everything this kernel reads from global memory comes through this thread's TMA loads: the one place that has
            // to wait for the stream predecessor (all other warps are ordered behind the loads through the mbarriers)
            if (prm.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
            int ia = 0, iw = 0;
            for (int w = pair_id; w < n_work; w += n_pairs) {
                const int n0 = (w % tiles_n) * BN + (int)rank * Cfg::kWRows;
                int pb, pi, pk;
                const int m0 = tile_m0(prm, w, tiles_n, (int)rank, pb, pi, pk);
                for (int g = 0; g < n_groups; ++g, ++ia) {
                    const int chunk = g / 3, kh = g - chunk * 3;
                    const int c0 = chunk * 64;
                    const int ea = ia % Cfg::kNA;
                    mbar_wait(&a_empty[ea], ((ia / Cfg::kNA) & 1) ^ 1);
                    uint8_t* ab = a_ring + ea * Cfg::kAEntry;
                    const uint32_t afull = mapa_u32(smem_u32(&a_full[ea]), 0);
                    const bool skip_a = (prm.dbg_flags & 1) && ia >= Cfg::kNA;
                    if (skip_a) {
                        if (rank == 0) mbar_arrive(&a_full[ea]);
                    } else {
                    if (rank == 0) mbar_arrive_expect_tx(&a_full[ea], 2 * Cfg::kABoxBytes * Cfg::kOperands);
                    const int arow = m0 + (kh - 1) * prm.Wp - 1;
                    tma_load_2d_pair(ab, &map_a_hi, afull, c0, arow);
                    if (PASSES >= 2) tma_load_2d_pair(ab + Cfg::kAPlane, &map_a_lo, afull, c0, arow);
                    }
                    for (int kw = 0; kw < 3; ++kw, ++iw) {
                        if (WRES && iw >= 9) continue;             // the nine taps are resident after the first tile
                        const int ew = iw % Cfg::kNW;
                        if (!WRES) mbar_wait(&w_empty[ew], ((iw / Cfg::kNW) & 1) ^ 1);
                        uint8_t* wb = w_ring + ew * Cfg::kWEntry;
                        const uint32_t wfull = mapa_u32(smem_u32(&w_full[ew]), 0);
                        if ((prm.dbg_flags & 1) && iw >= Cfg::kNW) {
                            if (rank == 0) mbar_arrive(&w_full[ew]);
                            continue;
                        }
                        if (rank == 0) mbar_arrive_expect_tx(&w_full[ew], 2 * Cfg::kWEntry);
                        const int kcol = (kh * 3 + kw) * prm.Cin + c0;
                        tma_load_2d_pair(wb, &map_w_hi, wfull, kcol, n0);
                        if (PASSES >= 2) tma_load_2d_pair(wb + Cfg::kWPlane, &map_w_lo, wfull, kcol, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            constexpr uint32_t idesc = PASSES == 2 ? make_idesc_f16(2 * kBM, BN) : make_idesc_bf16(2 * kBM, BN);
            constexpr uint32_t idesc8 = make_idesc_e5m2(2 * kBM, BN);
            (void)idesc8;
            int ia = 0, iw = 0, tl = 0;
            for (int w = pair_id; w < n_work; w += n_pairs, ++tl) {
                const int acc = tl & 1;
                const uint32_t d_tmem = tmem_base + acc * Cfg::kAccCols;
                mbar_wait(&tmem_empty[acc], ((tl >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int g = 0; g < n_groups; ++g, ++ia) {
                    const int ea = ia % Cfg::kNA;
                    mbar_wait(&a_full[ea], (ia / Cfg::kNA) & 1);
                    if (ia == 0 && lane == 0) stamp(prm, 2);   // first activation box landed
                    // descriptor low words: base of this A entry, advanced by whole rows (kw) and along K (k) with adds
                    const uint32_t da_base = kmajor_desc_lo(smem_u32(a_ring + ea * Cfg::kAEntry));
                    for (int kw = 0; kw < 3; ++kw, ++iw) {
                        const int ew = WRES ? g * 3 + kw : iw % Cfg::kNW;   // WRES: one 64-channel chunk, g = kh
                        mbar_wait(&w_full[ew], WRES ? 0u : (uint32_t)((iw / Cfg::kNW) & 1));
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t da_kw = da_base + kw * (128 >> 4);
                            const uint32_t db_base = kmajor_desc_lo(smem_u32(w_ring + ew * Cfg::kWEntry));
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t da = da_kw + k * (32 >> 4), db = db_base + k * (32 >> 4);
                                if (k < prm.k16_steps) {   // (channels >= cin_valid of a zero-padded chunk: nothing to add)
                                    mma_f16_pair_lo(d_tmem, da, db, idesc, (g > 0 || kw > 0 || k > 0) ? 1u : 0u);
                                    if (PASSES == 3) {
                                        mma_f16_pair_lo(d_tmem, da + (Cfg::kAPlane >> 4), db, idesc, 1u);
                                        mma_f16_pair_lo(d_tmem, da, db + (Cfg::kWPlane >> 4), idesc, 1u);
                                    }
                                }
                                if (PASSES == 2 && k < prm.e5_ksteps)
                                    mma_f8_pair_lo(d_tmem, da + (Cfg::kAPlane >> 4), db + (Cfg::kWPlane >> 4), idesc8, 1u);
                            }
                            if (!WRES) mma_commit_pair(&w_empty[ew], 3);
                            if (kw == 2) mma_commit_pair(&a_empty[ea], 3);
                            if (kw == 2 && g == n_groups - 1) mma_commit_pair(&tmem_full[acc], 3);
                        }
                        __syncwarp();
                    }
                }
            }
            if (lane == 0) stamp(prm, 3);   // last MMA issued
        }
    } else {
        const int q = warp & 3;
        int tl = 0;
        for (int w = pair_id; w < n_work; w += n_pairs, ++tl) {
            const int n0 = (w % tiles_n) * BN;
            int pb, pi, pk;
            const int m0 = tile_m0(prm, w, tiles_n, (int)rank, pb, pi, pk);
            if (w + n_pairs >= n_work && warp == 2 && lane == 0 && prm.stamps != nullptr && blockIdx.x == 0) {
                mbar_wait(&tmem_full[tl & 1], (tl >> 1) & 1);
                stamp(prm, 4);   // last accumulator complete (MMAs retired)
            }
            if constexpr (POOL)
                epilogue_tile_pool<BN, Cfg::kAccCols>(prm, tmem_base, tl & 1, pb, pi, pk, (int)rank, q, (warp - 2) >> 2, lane,
                                                      tmem_full, tmem_empty, tl, bias_s, xbuf, x_full, x_empty);
            else
                epilogue_tile<BN, Cfg::kAccCols, true, LEAN>(prm, tmem_base, tl & 1, m0, n0, q, (warp - 2) >> 2, lane, tmem_full, tmem_empty, tl, bias_s);
        }
        if (warp == 2 && lane == 0) stamp(prm, 5);   // this warp's epilogue done
    }
    tc_fence_before();
    // (execution-only barriers: a release arrive = MEMBAR.ALL.GPU would wait for the epilogue's global stores to land)
    cluster_sync_relaxed();  // the leader's MMAs read the peer's shared memory / write its TMEM until the last commit
    if (threadIdx.x == 0) stamp(prm, 6);   // all warps of both CTAs done
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    }
    cluster_sync_relaxed();
    if (threadIdx.x == 0) stamp(prm, 7);   // exit
}

// ------------------------------------------------------------------------------------------------
// fc layers over a few hundred ROIs (fc6: 300 x 25088 -> 2048): SWAPPED CTA-pair GEMM.  The layer is bound by weight
// bytes (fc6: 205 MB of bf16 hi/lo per branch).  With the ROIs on the M side (three 128-row tiles) every weight tile
// is pulled through L2 three times and each CTA moves as many activation bytes as weight bytes (1.2 GB L2->SM per
// fc6, the L2 ceiling).  Here the WEIGHTS are the M operand of a 256-row pair tile and ALL ROIs are the N operand
// (two MMA column chunks of C <= 256, the ROI rows split between the two CTAs), so each weight byte is fetched
// exactly once and the (small, L2-resident) activation matrix once per 256 output features: 0.44 GB per fc6.
//   D^T[f, r] = sum_k W[f, k] * A[r, k]        (both operands K-major, as they are stored)
// K is split over work items; partial sums go to the fp32 accumulator out[r, f] with red.global.add -- lanes are
// consecutive features, so the adds of one column are coalesced.  Same 3-pass bf16 hi/lo arithmetic.
// ------------------------------------------------------------------------------------------------
struct FcSwapParams {
    int F, R, K;            // output features (rows of W), ROIs (rows of A), reduction length (multiple of 64)
    int C;                  // columns per MMA chunk: 2 chunks cover R_pad = 2C ROIs; C % 16 == 0, C <= 256
    int k_steps_total, k_steps_per_split, n_split, n_work;
    int stages, stage_bytes;
    float* out;             // (R, ld) fp32, zeroed by the caller
    int ld;
    int passes;             // 3: bf16 hi/lo operands; 2: f16e5 operands (fp16 plane + e5m2 byte plane, see common.cuh)
    float out_scale;        // 2^-12 for f16e5 (the fp16 weight plane holds 4096 w), else 1
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
fc_swapped_pair_kernel(const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                       const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                       const FcSwapParams prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kWPlane = kBM * 128;                   // 128 weight rows x 64 channels
    const int half_rows = prm.C / 2;                     // ROI rows this CTA stages per chunk
    const int a_chunk = half_rows * 128;                 // bytes of one chunk tile
    const int a_plane = 2 * a_chunk;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + prm.stages * prm.stage_bytes);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* tmem_full = empty_bar + 8;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&map_w_hi);
        prefetch_tensormap(&map_w_lo);
        prefetch_tensormap(&map_a_hi);
        prefetch_tensormap(&map_a_lo);
        for (int i = 0; i < prm.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 2 * kEpiWarps);
        fence_barrier_init();
    }
    cluster_sync_all();                                  // see conv3x3_pair_kernel: both CTAs running before the alloc
    if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    if (warp == 1) tmem_relinquish_pair();
    const uint32_t tmem_base = *tmem_slot;
    const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int w = pair_id; w < prm.n_work; w += n_pairs) {
                const int mt = w / prm.n_split, sp = w - mt * prm.n_split;
                const int f0 = mt * (2 * kBM) + (int)rank * kBM;
                const int k_begin = sp * prm.k_steps_per_split;
                const int k_end = min(k_begin + prm.k_steps_per_split, prm.k_steps_total);
                for (int ks = k_begin; ks < k_end; ++ks, ++it) {
                    const int s = it % prm.stages;
                    mbar_wait(&empty_bar[s], ((it / prm.stages) & 1) ^ 1);
                    uint8_t* st = smem + s * prm.stage_bytes;
                    const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * prm.stage_bytes);
                    const int c0 = ks * 64;
                    tma_load_2d_pair(st, &map_w_hi, fb, c0, f0);
                    tma_load_2d_pair(st + kWPlane, &map_w_lo, fb, c0, f0);
                    uint8_t* ab = st + 2 * kWPlane;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r0 = h * prm.C + (int)rank * half_rows;
                        tma_load_2d_pair(ab + h * a_chunk, &map_a_hi, fb, c0, r0);
                        tma_load_2d_pair(ab + a_plane + h * a_chunk, &map_a_lo, fb, c0, r0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            const uint32_t idesc = prm.passes == 2 ? make_idesc_f16(2 * kBM, prm.C) : make_idesc_bf16(2 * kBM, prm.C);
            const uint32_t idesc8 = make_idesc_e5m2(2 * kBM, prm.C);
            int it = 0, tl = 0;
            for (int w = pair_id; w < prm.n_work; w += n_pairs, ++tl) {
                const int sp = w % prm.n_split;
                const int k_begin = sp * prm.k_steps_per_split;
                const int k_end = min(k_begin + prm.k_steps_per_split, prm.k_steps_total);
                mbar_wait(tmem_empty, (tl & 1) ^ 1);
                tc_fence_after();
                for (int ks = k_begin; ks < k_end; ++ks, ++it) {
                    const int s = it % prm.stages;
                    mbar_wait(&full_bar[s], (it / prm.stages) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t w_hi = smem_u32(smem + s * prm.stage_bytes);
                        const uint32_t w_lo = w_hi + kWPlane;
                        const uint32_t a_hi = w_lo + kWPlane;
                        const uint32_t a_lo = a_hi + a_plane;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint32_t d_tmem = tmem_base + h * prm.C;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t dw = make_kmajor_desc(w_hi + k * 32, 128);
                                const uint64_t dwl = make_kmajor_desc(w_lo + k * 32, 128);
                                const uint64_t da = make_kmajor_desc(a_hi + h * a_chunk + k * 32, 128);
                                const uint64_t dal = make_kmajor_desc(a_lo + h * a_chunk + k * 32, 128);
                                mma_bf16_ss_pair(d_tmem, dw, da, idesc, (ks > k_begin || k > 0) ? 1u : 0u);
                                if (prm.passes == 2) {   // [W_l8 | W_h8] . [A_h8 ; A_l8]: both correction terms, one e5m2 pass
                                    mma_f8_ss_pair(d_tmem, dwl, dal, idesc8, 1u);
                                } else {
                                    mma_bf16_ss_pair(d_tmem, dwl, da, idesc, 1u);
                                    mma_bf16_ss_pair(d_tmem, dw, dal, idesc, 1u);
                                }
                            }
                        }
                        mma_commit_pair(&empty_bar[s], 3);
                        if (ks == k_end - 1) mma_commit_pair(tmem_full, 3);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int n_chunks = (2 * prm.C) / 32;
        int tl = 0;
        for (int w = pair_id; w < prm.n_work; w += n_pairs, ++tl) {
            const int mt = w / prm.n_split;
            const int f = mt * (2 * kBM) + (int)rank * kBM + q * 32 + lane;   // this thread's output feature
            mbar_wait(tmem_full, tl & 1);
            tc_fence_after();
            const uint32_t taddr_row = tmem_base + (uint32_t(q * 32) << 16);
            for (int ci = half; ci < n_chunks; ci += 2) {
                uint32_t v[32];
                __syncwarp();
                tmem_ld_32x32(taddr_row + ci * 32, v);
                tmem_ld_wait();
                if (ci + 2 >= n_chunks) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(tmem_empty), 0));
                }
                if (f < prm.F) {
                    float* o = prm.out + (long long)(ci * 32) * prm.ld + f;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (ci * 32 + j < prm.R) red_add_f32(o + (long long)j * prm.ld, __uint_as_float(v[j]) * prm.out_scale);
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
    cluster_sync_all();
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int tap_reuse_mode() {  // MV3D_TAP_REUSE=0 selects the plain nine-box kernel (A/B comparisons); default on
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("MV3D_TAP_REUSE");
        mode = e ? atoi(e) : 1;
    }
    return mode;
}

static int gemm_dbg_flags() {   // MV3D_GEMM_DBG, see GemmParams::dbg_flags (8 = 128-bit instead of 256-bit epilogue stores)
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("MV3D_GEMM_DBG"); dbg = e ? atoi(e) : 0; }
    return dbg;
}
static long long* g_stamps = nullptr;   // mv3d_gemm_set_stamps
static int g_pair_mode = -1;
static int pair_mode() {  // MV3D_PAIR=0 selects the single-CTA kernels (A/B comparisons); default on
    if (g_pair_mode < 0) {
        const char* e = getenv("MV3D_PAIR");
        g_pair_mode = e ? atoi(e) : 1;
    }
    return g_pair_mode;
}

static int g_pdl = -1;
static int pdl_mode() {   // MV3D_PDL=1: pair kernels are launched as programmatic dependents of their stream predecessor
    if (g_pdl < 0) { const char* e = getenv("MV3D_PDL"); g_pdl = e ? atoi(e) : 0; }
    return g_pdl;
}
static int g_e5_ksteps = -1;
static int e5_ksteps_mode() {   // MV3D_F16E5_TERMS=1: experiment -- drop the activation-residual term (A_l W_h) of the f16e5 product
    if (g_e5_ksteps < 0) { const char* e = getenv("MV3D_F16E5_TERMS"); g_e5_ksteps = (e && atoi(e) == 1) ? 2 : 4; }
    return g_e5_ksteps;
}

template <int BN, int PASSES, bool LEAN, bool WRES = false, bool POOL = false>
static int launch_pair_impl(const mv3d_gemm_desc* d, cudaStream_t stream) {
    using Cfg = PairCfg<BN, PASSES, WRES, POOL>;
    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
    const uint64_t kcols = (uint64_t)9 * d->Cin;
    int rc;
    if ((rc = make_map_2d(&ma_hi, d->d_a_hi, d->M, d->Cin, kReuseRows, 64)) != MV3D_OK) return rc;
    if ((rc = make_map_2d(&mw_hi, d->d_w_hi, d->N, kcols, Cfg::kWRows, 64)) != MV3D_OK) return rc;
    if (PASSES >= 2) {
        if ((rc = make_map_2d(&ma_lo, d->d_a_lo, d->M, d->Cin, kReuseRows, 64)) != MV3D_OK) return rc;
        if ((rc = make_map_2d(&mw_lo, d->d_w_lo, d->N, kcols, Cfg::kWRows, 64)) != MV3D_OK) return rc;
    } else {
        ma_lo = ma_hi;
        mw_lo = mw_hi;
    }
    GemmParams p;
    p.M = d->M; p.N = d->N; p.Cin = d->Cin; p.taps = 9; p.Hp = d->Hp; p.Wp = d->Wp;
    p.H = d->Hp - 1; p.W = d->Wp - 1;
    p.k_chunks = d->Cin / 64;
    p.k_steps_total = 9 * p.k_chunks;
    p.k_steps_per_split = p.k_steps_total;
    p.split_k = 1;
    p.bias = d->d_bias; p.relu = d->relu;
    p.out_hi = static_cast<__nv_bfloat16*>(d->d_out_hi);
    p.out_lo = static_cast<__nv_bfloat16*>(d->d_out_lo);
    p.ld_out = d->ld_out;
    p.out_f32 = d->d_out_f32; p.ld_f32 = d->ld_f32; p.f32_dense = d->f32_dense;
    p.mask_hi = static_cast<const __nv_bfloat16*>(d->d_mask_hi); p.ld_mask = d->ld_mask; p.mask_scale = d->mask_scale;
    p.addend = d->d_addend_f32; p.ld_addend = d->ld_addend;
    p.out_fmt = d->out_fmt; p.acc_scale = (PASSES == 2) ? 1.f / kF16E5Scale : 1.f;
    p.dbg_flags = gemm_dbg_flags();
    p.stamps = g_stamps;
    p.e5_ksteps = e5_ksteps_mode();
    p.k16_steps = (d->cin_valid > 0 && d->Cin == 64) ? ceil_div(d->cin_valid < 64 ? d->cin_valid : 64, 16) : 4;
    p.softmax_cols = d->softmax_cols;
    p.pool = 0; p.pool_Ho = p.pool_Wo = p.pool_nblk = 0; p.pdl = 0;
    p.tiles_n = d->N / BN;
    p.tiles_m = ceil_div(d->M, 2 * kBM);
    p.n_work = p.tiles_n * p.tiles_m;
    if (POOL) {   // work items: (frame, row pair, 128-column block), one N tile
        p.pool = 1;
        p.pool_Ho = (d->Hp - 1) / 2; p.pool_Wo = (d->Wp - 1) / 2;
        p.pool_nblk = ceil_div(2 * p.pool_Wo, 128);
        p.tiles_n = 1;
        p.tiles_m = (d->M / (d->Hp * d->Wp)) * p.pool_Ho * p.pool_nblk;
        p.n_work = p.tiles_m;
    }
    auto kern = conv3x3_pair_kernel<BN, PASSES, LEAN, WRES, POOL>;
    static int max_pairs = 0;  // per instantiation
    if (max_pairs == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) { set_last_cuda_error(e); return MV3D_ERR_LAUNCH; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(num_sms() & ~1, 1, 1);
        cfg.blockDim = dim3(kGemmThreads, 1, 1);
        cfg.dynamicSmemBytes = Cfg::kSmemBytes;
        int n = 0;  // co-resident clusters of 2 (one CTA per SM): the persistent grid
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms() / 2; }
        max_pairs = n < num_sms() / 2 ? n : num_sms() / 2;
    }
    const int pairs = p.n_work < max_pairs ? p.n_work : max_pairs;
    if (pdl_mode()) {
        // programmatic dependent launch: this kernel's CTAs may be scheduled while the previous kernel of the stream drains
        // (each SM as soon as its CTA of that kernel has exited); the producer warp's griddepcontrol.wait holds every
        // read of the previous kernel's output until that kernel has completed
        p.pdl = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs, 1, 1);
        cfg.blockDim = dim3(kGemmThreads, 1, 1);
        cfg.dynamicSmemBytes = Cfg::kSmemBytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma_hi, ma_lo, mw_hi, mw_lo, p);
        if (e != cudaSuccess) { set_last_cuda_error(e); return MV3D_ERR_LAUNCH; }
        return MV3D_OK;
    }
    kern<<<2 * pairs, kGemmThreads, Cfg::kSmemBytes, stream>>>(ma_hi, ma_lo, mw_hi, mw_lo, p);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

// The forward-inference form of a descriptor (plain bias / ReLU epilogue, vector-store alignment) takes the LEAN kernel.
static bool lean_epilogue_ok(const mv3d_gemm_desc* d) {
    auto al32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; };
    if (d->split_k > 1 || d->d_mask_hi || d->d_addend_f32 || d->softmax_cols > 0 || (gemm_dbg_flags() & ~(4 | 16)) != 0) return false;
    if (d->d_out_hi && d->out_fmt == MV3D_FMT_BF16X2 && !(d->ld_out % 16 == 0 && al32(d->d_out_hi) && al32(d->d_out_lo))) return false;
    if (d->d_out_f32 && !(d->ld_f32 % 8 == 0 && al32(d->d_out_f32))) return false;
    return true;
}

static int g_wres_mode = -1;
static bool wres_mode() {   // MV3D_WRES=0: weights through the ring also in the 64 -> 64 layers (A/B comparisons)
    if (g_wres_mode < 0) { const char* e = getenv("MV3D_WRES"); g_wres_mode = e ? atoi(e) : 1; }
    return g_wres_mode != 0;
}

template <int BN, int PASSES>
static int launch_pair(const mv3d_gemm_desc* d, cudaStream_t stream) {
    if constexpr ((BN == 64 || BN == 128) && PASSES >= 2) {
        if (d->pool) {   // 2x2 max-pool fused into the epilogue (validated by mv3d_conv_gemm)
            if constexpr (BN == 64) {
                if (d->Cin == 64 && wres_mode()) return launch_pair_impl<BN, PASSES, true, true, true>(d, stream);
            }
            return launch_pair_impl<BN, PASSES, true, false, true>(d, stream);
        }
    }
    if constexpr (BN == 64 && PASSES >= 2) {
        if (d->Cin == 64 && wres_mode())
            return lean_epilogue_ok(d) ? launch_pair_impl<BN, PASSES, true, true>(d, stream)
                                       : launch_pair_impl<BN, PASSES, false, true>(d, stream);
    }
    return lean_epilogue_ok(d) ? launch_pair_impl<BN, PASSES, true>(d, stream) : launch_pair_impl<BN, PASSES, false>(d, stream);
}

// CTA-pair tiles for the tap-reuse 3x3 convs whose channel count tiles exactly; pair tile N = min(N, 256).
template <int PASSES>
static int try_launch_pair(const mv3d_gemm_desc* d, cudaStream_t stream, bool* taken) {
    *taken = true;
    if (d->N % 256 == 0) return launch_pair<256, PASSES>(d, stream);
    if (d->N == 128) return launch_pair<128, PASSES>(d, stream);
    if (d->N == 64) return launch_pair<64, PASSES>(d, stream);
    *taken = false;
    return MV3D_OK;
}

// fc over <= 512 ROIs with a split-K fp32 accumulator: the swapped CTA-pair kernel (weights stream through once).
static bool fc_swap_applicable(const mv3d_gemm_desc* d) {
    return d->taps == 1 && d->Hp == 0 && (d->passes == 3 || d->passes == 2) && d->split_k > 1 && d->d_out_f32 && d->Cin % 64 == 0 &&
           d->N % 256 == 0 && d->M >= 64 && d->M <= 512 && pair_mode() != 0;
}

static int launch_fc_swapped(const mv3d_gemm_desc* d, cudaStream_t stream) {
    FcSwapParams p;
    p.F = d->N; p.R = d->M; p.K = d->Cin;
    p.C = ceil_div(d->M, 32) * 16;
    p.k_steps_total = d->Cin / 64;
    const int m_tiles = d->N / (2 * kBM);
    const int pairs_avail = num_sms() / 2;
    int split = pairs_avail / m_tiles;                    // one wave of work items
    if (split < 1) split = 1;
    if (split > p.k_steps_total) split = p.k_steps_total;
    p.k_steps_per_split = ceil_div(p.k_steps_total, split);
    p.n_split = ceil_div(p.k_steps_total, p.k_steps_per_split);
    p.n_work = m_tiles * p.n_split;
    p.stage_bytes = 2 * kBM * 128 + 2 * (2 * (p.C / 2) * 128);
    p.stages = (220 * 1024) / p.stage_bytes;
    if (p.stages > 8) p.stages = 8;
    if (p.stages < 2) return MV3D_ERR_ARG;
    p.out = d->d_out_f32; p.ld = d->ld_f32;
    p.passes = d->passes; p.out_scale = d->passes == 2 ? 1.f / kF16E5Scale : 1.f;
    CUtensorMap mw_hi, mw_lo, ma_hi, ma_lo;
    int rc;
    if ((rc = make_map_2d(&mw_hi, d->d_w_hi, d->N, d->Cin, kBM, 64)) != MV3D_OK) return rc;
    if ((rc = make_map_2d(&mw_lo, d->d_w_lo, d->N, d->Cin, kBM, 64)) != MV3D_OK) return rc;
    if ((rc = make_map_2d(&ma_hi, d->d_a_hi, d->M, d->Cin, p.C / 2, 64)) != MV3D_OK) return rc;
    if ((rc = make_map_2d(&ma_lo, d->d_a_lo, d->M, d->Cin, p.C / 2, 64)) != MV3D_OK) return rc;
    const int smem_bytes = p.stages * p.stage_bytes + 1024 + 256;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(fc_swapped_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_last_cuda_error(e); return MV3D_ERR_LAUNCH; }
        attr_set = true;
    }
    const int pairs = p.n_work < pairs_avail ? p.n_work : pairs_avail;
    fc_swapped_pair_kernel<<<2 * pairs, kGemmThreads, smem_bytes, stream>>>(mw_hi, mw_lo, ma_hi, ma_lo, p);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

template <int BN, int PASSES>
static int launch_reuse(const mv3d_gemm_desc* d, cudaStream_t stream) {
    using Cfg = ReuseCfg<BN, PASSES>;
    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
    const uint64_t kcols = (uint64_t)9 * d->Cin;
    int rc;
    if ((rc = make_map_2d(&ma_hi, d->d_a_hi, d->M, d->Cin, kReuseRows, 64)) != MV3D_OK) return rc;
    if ((rc = make_map_2d(&mw_hi, d->d_w_hi, d->N, kcols, BN, 64)) != MV3D_OK) return rc;
    if (PASSES >= 2) {
        if ((rc = make_map_2d(&ma_lo, d->d_a_lo, d->M, d->Cin, kReuseRows, 64)) != MV3D_OK) return rc;
        if ((rc = make_map_2d(&mw_lo, d->d_w_lo, d->N, kcols, BN, 64)) != MV3D_OK) return rc;
    } else {
        ma_lo = ma_hi;
        mw_lo = mw_hi;
    }
    GemmParams p;
    p.M = d->M; p.N = d->N; p.Cin = d->Cin; p.taps = 9; p.Hp = d->Hp; p.Wp = d->Wp;
    p.H = d->Hp - 1; p.W = d->Wp - 1;
    p.k_chunks = d->Cin / 64;
    p.k_steps_total = 9 * p.k_chunks;
    p.k_steps_per_split = p.k_steps_total;
    p.split_k = 1;
    p.bias = d->d_bias; p.relu = d->relu;
    p.out_hi = static_cast<__nv_bfloat16*>(d->d_out_hi);
    p.out_lo = static_cast<__nv_bfloat16*>(d->d_out_lo);
    p.ld_out = d->ld_out;
    p.out_f32 = d->d_out_f32; p.ld_f32 = d->ld_f32; p.f32_dense = d->f32_dense;
    p.mask_hi = static_cast<const __nv_bfloat16*>(d->d_mask_hi); p.ld_mask = d->ld_mask; p.mask_scale = d->mask_scale;
    p.addend = d->d_addend_f32; p.ld_addend = d->ld_addend;
    p.out_fmt = d->out_fmt; p.acc_scale = (PASSES == 2) ? 1.f / kF16E5Scale : 1.f;
    p.dbg_flags = gemm_dbg_flags();
    p.stamps = nullptr;
    p.softmax_cols = d->softmax_cols;
    p.pool = 0; p.pool_Ho = p.pool_Wo = p.pool_nblk = 0; p.pdl = 0;
    p.tiles_n = ceil_div(d->N, BN);
    p.tiles_m = ceil_div(d->M, kBM);
    p.n_work = p.tiles_n * p.tiles_m;
    auto kern = conv3x3_reuse_kernel<BN, PASSES>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) { set_last_cuda_error(e); return MV3D_ERR_LAUNCH; }
        attr_set = true;
    }
    const int grid = p.n_work < num_sms() ? p.n_work : num_sms();
    kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ma_hi, ma_lo, mw_hi, mw_lo, p);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

template <int BN, int KC, int PASSES>
static int launch_gemm(const mv3d_gemm_desc* d, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, KC, PASSES>;
    if (KC == 64 && d->taps == 9 && d->split_k <= 1 && tap_reuse_mode() != 0) return launch_reuse<BN, PASSES>(d, stream);
    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
    const uint64_t kcols = (uint64_t)d->taps * d->Cin;
    int rc;
    if ((rc = make_map_2d(&ma_hi, d->d_a_hi, d->M, d->Cin, kBM, KC)) != MV3D_OK) return rc;
    if ((rc = make_map_2d(&mw_hi, d->d_w_hi, d->N, kcols, BN, KC)) != MV3D_OK) return rc;
    if (PASSES == 3) {
        if ((rc = make_map_2d(&ma_lo, d->d_a_lo, d->M, d->Cin, kBM, KC)) != MV3D_OK) return rc;
        if ((rc = make_map_2d(&mw_lo, d->d_w_lo, d->N, kcols, BN, KC)) != MV3D_OK) return rc;
    } else {
        ma_lo = ma_hi;
        mw_lo = mw_hi;
    }
    GemmParams p;
    p.M = d->M; p.N = d->N; p.Cin = d->Cin; p.taps = d->taps; p.Hp = d->Hp; p.Wp = d->Wp;
    p.H = d->Hp - 1; p.W = d->Wp - 1;
    p.k_chunks = d->Cin / KC;
    p.k_steps_total = d->taps * p.k_chunks;
    int split = d->split_k > 1 ? d->split_k : 1;
    if (split > p.k_steps_total) split = p.k_steps_total;
    p.k_steps_per_split = ceil_div(p.k_steps_total, split);
    split = ceil_div(p.k_steps_total, p.k_steps_per_split);  // no empty z-slices
    p.split_k = split;
    p.bias = d->d_bias; p.relu = d->relu;
    p.out_hi = static_cast<__nv_bfloat16*>(d->d_out_hi);
    p.out_lo = static_cast<__nv_bfloat16*>(d->d_out_lo);
    p.ld_out = d->ld_out;
    p.out_f32 = d->d_out_f32; p.ld_f32 = d->ld_f32; p.f32_dense = d->f32_dense;
    p.mask_hi = static_cast<const __nv_bfloat16*>(d->d_mask_hi); p.ld_mask = d->ld_mask; p.mask_scale = d->mask_scale;
    p.addend = d->d_addend_f32; p.ld_addend = d->ld_addend;
    p.out_fmt = d->out_fmt; p.acc_scale = (PASSES == 2) ? 1.f / kF16E5Scale : 1.f;
    p.dbg_flags = gemm_dbg_flags();
    p.stamps = nullptr;
    p.softmax_cols = d->softmax_cols;
    p.pool = 0; p.pool_Ho = p.pool_Wo = p.pool_nblk = 0; p.pdl = 0;

    auto kern = conv_gemm_kernel<BN, KC, PASSES>;
    static bool attr_set = false;  // per instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) { set_last_cuda_error(e); return MV3D_ERR_LAUNCH; }
        attr_set = true;
    }
    p.tiles_n = ceil_div(d->N, BN);
    p.tiles_m = ceil_div(d->M, kBM);
    const long long n_work = (long long)p.tiles_n * p.tiles_m * split;
    if (n_work > 0x7fffffffLL) return MV3D_ERR_ARG;
    p.n_work = (int)n_work;
    const int grid = p.n_work < num_sms() ? p.n_work : num_sms();
    kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ma_hi, ma_lo, mw_hi, mw_lo, p);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

template <int KC, int PASSES>
static int dispatch_bn(const mv3d_gemm_desc* d, cudaStream_t s) {
    const int n = d->N;
    if constexpr (KC == 64) {
        if (d->taps == 9 && d->split_k <= 1 && tap_reuse_mode() != 0 && pair_mode() != 0) {
            bool taken = false;
            const int rc = try_launch_pair<PASSES>(d, s, &taken);
            if (taken) return rc;
        }
    }
    // widest tile that the operand staging affords: 256 columns single pass, 128 in the 3-pass mode
    if constexpr (PASSES == 1) {
        if (n > 128) return launch_gemm<256, KC, PASSES>(d, s);
    }
    if (n > 64) return launch_gemm<128, KC, PASSES>(d, s);
    if (n > 32) return launch_gemm<64, KC, PASSES>(d, s);
    return launch_gemm<32, KC, PASSES>(d, s);
}

}  // namespace mv3d

extern "C" __attribute__((visibility("default"))) int mv3d_gemm_set_pair_mode(int on) {
    const int prev = mv3d::pair_mode();
    mv3d::g_pair_mode = on ? 1 : 0;
    return prev;
}

// Measurement only: device buffer of 8 int64 that pair 0 of every conv3x3_pair_kernel launch fills with clock64() at its
// phase boundaries (0 start, 1 set-up done, 2 first operands landed, 3 last MMA issued, 4 last accumulator complete,
// 5 epilogue done, 6 both CTAs done, 7 exit); nullptr switches it off.  tools/gemm_phases.py prints the breakdown.
extern "C" __attribute__((visibility("default"))) int mv3d_gemm_set_stamps(void* d_stamps) {
    mv3d::g_stamps = static_cast<long long*>(d_stamps);
    return MV3D_OK;
}

extern "C" __attribute__((visibility("default"))) int mv3d_conv_gemm(const mv3d_gemm_desc* d, void* stream) {
    using namespace mv3d;
    MV3D_REQUIRE(d != nullptr && d->M > 0 && d->N > 0 && d->Cin > 0);
    MV3D_REQUIRE(d->taps == 1 || d->taps == 9);
    MV3D_REQUIRE(d->Cin % 16 == 0);
    MV3D_REQUIRE(d->passes == 1 || d->passes == 2 || d->passes == 3);
    MV3D_REQUIRE(d->d_a_hi && d->d_w_hi);
    MV3D_REQUIRE(d->passes == 1 || (d->d_a_lo && d->d_w_lo));
    MV3D_REQUIRE(d->out_fmt == MV3D_FMT_BF16X2 || d->out_fmt == MV3D_FMT_F16E5);
    // f16e5 output: whole 64-channel chunks, both planes
    MV3D_REQUIRE(d->out_fmt != MV3D_FMT_F16E5 || !d->d_out_hi || (d->d_out_lo && d->N % 64 == 0 && d->ld_out % 64 == 0 && d->split_k <= 1 &&
                                                                   ((reinterpret_cast<uintptr_t>(d->d_out_hi) | reinterpret_cast<uintptr_t>(d->d_out_lo)) & 31) == 0));
    // f16e5 operands: the tap-reuse 3x3 kernels, and the swapped split-K fc kernel
    MV3D_REQUIRE(d->passes != 2 || (d->taps == 9 && d->Cin % 64 == 0 && d->split_k <= 1) || fc_swap_applicable(d));
    MV3D_REQUIRE(d->taps == 1 || (d->Hp > 1 && d->Wp > 1));
    MV3D_REQUIRE((d->Hp > 0) == (d->Wp > 0));
    MV3D_REQUIRE(d->Hp == 0 || d->M % (d->Hp * d->Wp) == 0);
    MV3D_REQUIRE(d->d_out_hi || d->d_out_f32);
    MV3D_REQUIRE(d->split_k <= 1 || d->d_out_f32);
    MV3D_REQUIRE(!d->f32_dense || d->Hp > 0);
    MV3D_REQUIRE(d->split_k <= 1 || (!d->d_mask_hi && !d->d_addend_f32));
    MV3D_REQUIRE(d->softmax_cols >= 0 && d->softmax_cols <= d->N && d->softmax_cols % 2 == 0);
    MV3D_REQUIRE(d->cin_valid >= 0 && d->cin_valid <= d->Cin);
    // fused 2x2 max-pool: CTA-pair tap-reuse kernel, one N tile of 64 / 128 channels, operand output only, plain epilogue
    MV3D_REQUIRE(!d->pool || (d->taps == 9 && (d->N == 64 || d->N == 128) && d->Cin % 64 == 0 && d->passes >= 2 && d->d_out_hi &&
                              !d->d_out_f32 && d->split_k <= 1 && d->Hp > 2 && d->Wp > 2 && pair_mode() != 0 &&
                              lean_epilogue_ok(d) && (d->out_fmt == MV3D_FMT_F16E5 || d->ld_out % 16 == 0)));
    MV3D_REQUIRE(d->softmax_cols == 0 || (d->d_out_f32 && !d->d_out_hi && !d->relu && d->split_k <= 1 && !d->d_mask_hi && !d->d_addend_f32));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (fc_swap_applicable(d)) return launch_fc_swapped(d, s);
    if (d->passes == 2) {
        if (pair_mode() != 0) {
            bool taken = false;
            const int rc = try_launch_pair<2>(d, s, &taken);
            if (taken) return rc;
        }
        return d->N > 64 ? launch_reuse<128, 2>(d, s) : launch_reuse<64, 2>(d, s);
    }
    const bool kc64 = (d->Cin % 64 == 0);
    const bool kc32 = !kc64 && (d->Cin % 32 == 0);   // K step 32 -> SWIZZLE_64B boxes (the im2col'd first layers, K = 32)
    if (d->passes == 3) return kc64 ? dispatch_bn<64, 3>(d, s) : (kc32 ? dispatch_bn<32, 3>(d, s) : dispatch_bn<16, 3>(d, s));
    return kc64 ? dispatch_bn<64, 1>(d, s) : (kc32 ? dispatch_bn<32, 1>(d, s) : dispatch_bn<16, 1>(d, s));
}
