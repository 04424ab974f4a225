// Geometry helpers shared by the proposal layer, the proposal-target layer and the fused multi-view ROI pool: the
// reference's BEV box (lib/utils/transform.py:113-142, numpy `//`), its 8-corner image projection
// (transform.py:290-315,369-386,483-500) with numpy's dtype pipeline, and this project's front-view box.
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
// One corner c (0..7) of the box in the reference's order -> its image coordinates (u, v) in float64.
__device__ __forceinline__ void img_corner_uv(const float* M, float xp, float xm, float yp, float ym, float zp, float zm,
                                              int c, double& u, double& v) {
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
    u = r[0] / r[2];
    v = r[1] / r[2];
}

// min / max over the 8 corners' (u, v) in corner order with numpy's NaN propagation, int32 cast with x86 semantics.
__device__ __forceinline__ void img_box_from_uv(const double* u8, const double* v8, int* img) {
    double umin = 0, umax = 0, vmin = 0, vmax = 0;
    bool nan_u = false, nan_v = false;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const double u = u8[c], v = v8[c];
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

__device__ __forceinline__ void corners_to_img_box(const float* M, float xp, float xm, float yp, float ym, float zp,
                                                   float zm, int* img) {
    double u8[8], v8[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) img_corner_uv(M, xp, xm, yp, ym, zp, zm, c, u8[c], v8[c]);
    img_box_from_uv(u8, v8, img);
}

// numpy float64 floor_divide == npy_divmod (numpy/core/src/npymath/npy_math_internal.h): fmod based.
__device__ __forceinline__ double npy_floor_divide(double a, double b) {
    if (b == 0.0) return a / b;
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0.0) {
        if ((b < 0) != (mod < 0)) div -= 1.0;
    }
    double floordiv;
    if (div != 0.0) {
        floordiv = floor(div);
        if (div - floordiv > 0.5) floordiv += 1.0;
    } else {
        floordiv = copysign(0.0, a / b);
    }
    return floordiv;
}

// np.maximum(np.minimum(v, hi), 0) with numpy's NaN propagation.
__device__ __forceinline__ float clip_np(float v, float hi) {
    if (v != v) return v;
    v = v < hi ? v : hi;
    return v > 0.f ? v : 0.f;
}

// BEV grid constants of transform.py:3-20 + the clip bounds of clip_boxes (im_w - 1, im_h - 1 as float32).
struct BevGrid {
    double xn, yn, x_min, y_min, res;
    float clip_x, clip_y;
};

// Half-extent sums of a 3-D box [x,y,z,l,w,h] (float32, as lidar_3d_to_bv / lidar_3d_to_corners form them).
struct BoxExtents {
    float xp, xm, yp, ym, zp, zm;
};
__device__ __forceinline__ BoxExtents box_extents(float px, float py, float pz, float pl, float pw, float ph) {
    const float hl = __fmul_rn(pl, 0.5f), hw = __fmul_rn(pw, 0.5f), hh = __fmul_rn(ph, 0.5f);
    BoxExtents e;
    e.xp = __fadd_rn(px, hl); e.xm = __fsub_rn(px, hl);
    e.yp = __fadd_rn(py, hw); e.ym = __fsub_rn(py, hw);
    e.zp = __fadd_rn(pz, hh); e.zm = __fsub_rn(pz, hh);
    return e;
}

// lidar_3d_to_bv (transform.py:132-140: f32 sums widened to f64, numpy `//`) + clip_boxes (bbox_transform.py:178-191)
// -> (x1, y1, x2, y2) integral-valued float32.
// coordinate q of the box: 0 = x1 (from y + w/2), 1 = y1 (from x + l/2), 2 = x2 (y - w/2), 3 = y2 (x - l/2)
__device__ __forceinline__ float bev_coord(const BevGrid& g, const BoxExtents& e, int q) {
    const bool is_x = (q & 1) == 0;
    const float src = q == 0 ? e.yp : (q == 1 ? e.xp : (q == 2 ? e.ym : e.xm));
    const float raw = is_x ? (float)(g.yn - npy_floor_divide((double)src - g.y_min, g.res))
                           : (float)(g.xn - npy_floor_divide((double)src - g.x_min, g.res));
    return clip_np(raw, is_x ? g.clip_x : g.clip_y);
}
__device__ __forceinline__ void extents_to_bev_box(const BevGrid& g, const BoxExtents& e, float* bv) {
#pragma unroll
    for (int q = 0; q < 4; ++q) bv[q] = bev_coord(g, e, q);
}

// Front view (no reference counterpart; DESIGN.md 'Front view'): cylindrical map geometry.
struct FvGeom {
    int H, W;
    double theta_min, dtheta, phi_max, dphi;  // radians
};

__device__ __forceinline__ void fv_coords(const FvGeom& g, double x, double y, double z, double& col, double& row) {
    col = (atan2(y, x) - g.theta_min) / g.dtheta;
    row = (g.phi_max - atan2(z, sqrt(x * x + y * y))) / g.dphi;
}

// FV rectangle of a 3-D box: floor of the FV coordinates of its 8 corners, min/max, clamped to the map
// -> [col_min, row_min, col_max, row_max]; a non-finite corner gives the empty rectangle (0,0,0,0).
// floor'd FV coordinates of corner k (0..7)
__device__ __forceinline__ void fv_corner(const FvGeom& g, const BoxExtents& e, int k, double& col, double& row) {
    const float x = (k & 1) ? e.xm : e.xp, y = ((k >> 1) & 1) ? e.ym : e.yp, z = (k >> 2) ? e.zm : e.zp;
    fv_coords(g, (double)x, (double)y, (double)z, col, row);
    col = floor(col);
    row = floor(row);
}
__device__ __forceinline__ void fv_box_from_corners(const FvGeom& g, const double* col8, const double* row8, float* o) {
    double cmin = 0, cmax = 0, rmin = 0, rmax = 0;
    bool bad = false;
    for (int k = 0; k < 8; ++k) {
        const double col = col8[k], row = row8[k];
        if (!(isfinite(col) && isfinite(row))) bad = true;
        if (k == 0) { cmin = cmax = col; rmin = rmax = row; }
        else {
            cmin = fmin(cmin, col); cmax = fmax(cmax, col);
            rmin = fmin(rmin, row); rmax = fmax(rmax, row);
        }
    }
    if (bad) { o[0] = o[1] = o[2] = o[3] = 0.f; return; }
    const double wmax = g.W - 1, hmax = g.H - 1;
    o[0] = (float)fmin(fmax(cmin, 0.0), wmax);
    o[1] = (float)fmin(fmax(rmin, 0.0), hmax);
    o[2] = (float)fmin(fmax(cmax, 0.0), wmax);
    o[3] = (float)fmin(fmax(rmax, 0.0), hmax);
}
__device__ __forceinline__ void extents_to_fv_box(const FvGeom& g, const BoxExtents& e, float* o) {
    double col8[8], row8[8];
    for (int k = 0; k < 8; ++k) fv_corner(g, e, k, col8[k], row8[k]);
    fv_box_from_corners(g, col8, row8, o);
}

}  // namespace mv3d
