// Status strings / error bookkeeping of the C ABI.
#include "common.cuh"

namespace mv3d {
static thread_local cudaError_t g_last_err = cudaSuccess;
void set_last_cuda_error(cudaError_t e) { g_last_err = e; }
}  // namespace mv3d

extern "C" __attribute__((visibility("default"))) int mv3d_version(void) { return 100; }

extern "C" __attribute__((visibility("default"))) const char* mv3d_status_string(int s) {
    switch (s) {
        case MV3D_OK: return "ok";
        case MV3D_ERR_ARG: return "invalid argument";
        case MV3D_ERR_WORKSPACE: return "workspace too small";
        case MV3D_ERR_LAUNCH: return "CUDA launch/runtime failure";
        case MV3D_ERR_DRIVER: return "CUDA driver entry point unavailable (cuTensorMapEncodeTiled)";
        default: return "unknown status";
    }
}
extern "C" __attribute__((visibility("default"))) int mv3d_last_cuda_error(void) { return (int)mv3d::g_last_err; }
extern "C" __attribute__((visibility("default"))) const char* mv3d_last_cuda_error_string(void) { return cudaGetErrorString(mv3d::g_last_err); }
