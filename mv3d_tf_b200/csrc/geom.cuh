// Geometry helpers shared by the proposal layer and the proposal-target layer: the reference's 8-corner image
// projection (lib/utils/transform.py:290-315,369-386,483-500) with numpy's dtype pipeline.
#pragma once
#include <limits.h>

#include "common.cuh"

namespace mv3d {

// C cast double -> int32 as x86 cvttsd2si does it (numpy astype(int32)): out of range / NaN -> INT_MIN.
__device__ __forceinline__ int cast_i32_x86(double v) {
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
    return (int)v;
}


// lidar_3d_to_corners (transform.py:305-313) + lidar_cnr_to_img (:483-500, :369-386): corners from the float32
// half-extent sums (xp = x + l/2, xm = x - l/2, ...), projected with the float32 3x4 matrix M in float64, divided by
// the third row (no abs), min/max over the 8 corners, cast to int32 with x86 semantics -> [xmin, ymin, xmax, ymax].
__device__ __forceinline__ void corners_to_img_box(const float* M, float xp, float xm, float yp, float ym, float zp,
                                                   float zm, int* img) {
    double umin = 0, umax = 0, vmin = 0, vmax = 0;
    bool nan_u = false, nan_v = false;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const bool sx = (c == 0 || c == 1 || c == 4 || c == 5);
        const bool sy = (c == 0 || c == 3 || c == 4 || c == 7);
        const bool sz = (c >= 4);
        const double X = sx ? xp : xm, Y = sy ? yp : ym, Z = sz ? zp : zm;
        double r[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            double acc = __dmul_rn((double)M[4 * q], X);
            acc = __dadd_rn(acc, __dmul_rn((double)M[4 * q + 1], Y));
            acc = __dadd_rn(acc, __dmul_rn((double)M[4 * q + 2], Z));
            acc = __dadd_rn(acc, __dmul_rn((double)M[4 * q + 3], 0.0));
            r[q] = acc;
        }
        const double u = r[0] / r[2], v = r[1] / r[2];
        nan_u |= (u != u);
        nan_v |= (v != v);
        if (c == 0) { umin = umax = u; vmin = vmax = v; }
        else {
            umin = u < umin ? u : umin; umax = u > umax ? u : umax;
            vmin = v < vmin ? v : vmin; vmax = v > vmax ? v : vmax;
        }
    }
    if (nan_u) umin = umax = nan("");
    if (nan_v) vmin = vmax = nan("");
    img[0] = cast_i32_x86(umin); img[1] = cast_i32_x86(vmin);
    img[2] = cast_i32_x86(umax); img[3] = cast_i32_x86(vmax);
}

}  // namespace mv3d
