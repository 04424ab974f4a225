// Front-view (FV) branch: cylindrical LiDAR raster and the FV region of interest of a 3-D proposal.
// The reference has NO front view (lib/networks/network.py:313-315 returns None for target='fv'); the semantics
// implemented here are this project's own specification (DESIGN.md, section 'Front view'; the CPU checker restates
// it as point_cloud_2_front / lidar_3d_to_fv) after the MV3D paper the reference's README.md:5 links.  All index math is
// float64 so that the device and the numpy specification agree bit for bit away from measure-zero cell edges.
#include "common.cuh"
#include "geom.cuh"

namespace mv3d {

// pass 1: last writer per cell (largest point index), table pre-zeroed
__global__ void fv_winner_kernel(const float* __restrict__ pts, int n, int stride, FvGeom g, int* __restrict__ table) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* p = pts + (size_t)i * stride;
        const double x = p[0], y = p[1], z = p[2];
        if (!(x > 0.0)) continue;
        double col, row;
        fv_coords(g, x, y, z, col, row);
        const double c = floor(col), r = floor(row);
        if (!(c >= 0.0 && c < (double)g.W && r >= 0.0 && r < (double)g.H)) continue;
        atomicMax(&table[(int)r * g.W + (int)c], i + 1);
    }
}

// pass 2: one thread per output cell (PAD grid incl. halo when pad != 0): channels [z, distance, reflectance]
__global__ void fv_write_kernel(const float* __restrict__ pts, int stride, FvGeom g, const int* __restrict__ table,
                                float* __restrict__ top, __nv_bfloat16* __restrict__ pad_hi,
                                __nv_bfloat16* __restrict__ pad_lo, int c_pad) {
    const int pad = pad_hi ? 1 : 0;
    const int Hout = g.H + pad, Wout = g.W + pad;
    const int total = Hout * Wout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ro = i / Wout, co = i - ro * Wout;
        const int r = ro, c = co - pad;
        float v[3] = {0.f, 0.f, 0.f};
        if (r < g.H && c >= 0) {
            const int w = table[r * g.W + c];
            if (w) {
                const float* p = pts + (size_t)(w - 1) * stride;
                const double x = p[0], y = p[1], z = p[2];
                v[0] = p[2];
                v[1] = (float)sqrt(x * x + y * y + z * z);
                v[2] = p[3];
            }
        }
        if (pad) {
            for (int ch = 0; ch < c_pad; ++ch) {
                __nv_bfloat16 h, l;
                split_bf16(ch < 3 ? v[ch] : 0.f, h, l);
                pad_hi[(size_t)i * c_pad + ch] = h;
                if (pad_lo) pad_lo[(size_t)i * c_pad + ch] = l;
            }
        }
        if (top && r < g.H && c >= 0) {
            float* o = top + ((size_t)r * g.W + c) * 3;
            o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
        }
    }
}

// rois_3d (R,7) [batch,x,y,z,l,w,h] -> rois_fv (R,5) [batch, col_min, row_min, col_max, row_max] (floor, clamped)
__global__ void rois_to_fv_kernel(const float* __restrict__ rois_3d, int R, const int* __restrict__ num_valid, FvGeom g,
                                  float* __restrict__ rois_fv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    float* o = rois_fv + (size_t)i * 5;
    if (num_valid && i >= *num_valid) {
        o[0] = o[1] = o[2] = o[3] = o[4] = 0.f;
        return;
    }
    const float* p = rois_3d + (size_t)i * 7;
    // corners exactly as lidar_3d_to_corners (transform.py:305-313): float32 half extents added to the centre
    o[0] = p[0];
    extents_to_fv_box(g, box_extents(p[1], p[2], p[3], p[4], p[5], p[6]), o + 1);
}

}  // namespace mv3d

using namespace mv3d;
#define MV3D_API extern "C" __attribute__((visibility("default")))

MV3D_API size_t mv3d_fv_raster_workspace_bytes(int H, int W) { return sizeof(int) * (size_t)H * (size_t)W; }

MV3D_API int mv3d_fv_raster(const float* d_points, int n_points, int point_stride, int H, int W, double theta_min_rad,
                            double dtheta_rad, double phi_max_rad, double dphi_rad, float* d_top, void* d_pad_hi,
                            void* d_pad_lo, int c_pad, void* d_workspace, size_t workspace_bytes, void* stream) {
    MV3D_REQUIRE(H > 0 && W > 0 && n_points >= 0 && point_stride >= 4 && (d_top || d_pad_hi));
    MV3D_REQUIRE(n_points == 0 || d_points);
    MV3D_REQUIRE(!d_pad_hi || c_pad >= 3);
    MV3D_REQUIRE(dtheta_rad > 0 && dphi_rad > 0);
    if (!d_workspace || workspace_bytes < mv3d_fv_raster_workspace_bytes(H, W)) return MV3D_ERR_WORKSPACE;
    FvGeom g{H, W, theta_min_rad, dtheta_rad, phi_max_rad, dphi_rad};
    cudaStream_t s = (cudaStream_t)stream;
    int* table = static_cast<int*>(d_workspace);
    cudaError_t e = cudaMemsetAsync(table, 0, sizeof(int) * (size_t)H * W, s);
    if (e != cudaSuccess) { set_last_cuda_error(e); return MV3D_ERR_LAUNCH; }
    if (n_points > 0) {
        int grid = ceil_div(n_points, 256);
        if (grid > 148 * 8) grid = 148 * 8;
        fv_winner_kernel<<<grid, 256, 0, s>>>(d_points, n_points, point_stride, g, table);
    }
    const int total = (H + (d_pad_hi ? 1 : 0)) * (W + (d_pad_hi ? 1 : 0));
    fv_write_kernel<<<ceil_div(total, 128), 128, 0, s>>>(d_points, point_stride, g, table, d_top,
                                                         (__nv_bfloat16*)d_pad_hi, (__nv_bfloat16*)d_pad_lo, c_pad);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}

MV3D_API int mv3d_rois_to_fv(const float* d_rois_3d, int R, const int* d_num_valid, int H, int W, double theta_min_rad,
                             double dtheta_rad, double phi_max_rad, double dphi_rad, float* d_rois_fv, void* stream) {
    MV3D_REQUIRE(R >= 0 && H > 0 && W > 0 && d_rois_fv && (R == 0 || d_rois_3d));
    if (R == 0) return MV3D_OK;
    FvGeom g{H, W, theta_min_rad, dtheta_rad, phi_max_rad, dphi_rad};
    rois_to_fv_kernel<<<ceil_div(R, 128), 128, 0, (cudaStream_t)stream>>>(d_rois_3d, R, d_num_valid, g, d_rois_fv);
    MV3D_CHECK_LAUNCH();
    return MV3D_OK;
}
