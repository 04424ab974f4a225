// Host-side TMA helpers shared by the tcgen05 GEMM kernels (tensor-map encoding through the driver entry point;
// no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace mv3d {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D bf16 row-major (rows, cols) matrix, box (box_rows, box_cols), swizzle matching box_cols*2 bytes.
static inline int make_map_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                              uint32_t box_cols, uint64_t ld = 0) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return MV3D_ERR_DRIVER;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {(ld ? ld : cols) * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : box_cols * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                 : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? MV3D_OK : MV3D_ERR_DRIVER;
}

static inline int num_sms() {
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    }
    return n_sm;
}


}  // namespace mv3d
