/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the scalar loops on the MV3D hot
 * path.  Never linked into the product library; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs load it.
 *
 * Every function cites the reference lines (relative to /root/reference) it restates.
 * Compile: gcc -O2 -fPIC -shared -ffp-contract=off -fno-fast-math (see oracle/build.py).
 * -ffp-contract=off matters: the reference arithmetic is separate IEEE mul/add roundings.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

/* lib/nms/cpu_nms.pyx:17-68 (rule_ge=1: `ovr >= thresh` in double, thresh a Python float)
 * lib/nms/nms_kernel.cu:24-32,71 (rule_ge=0: `ovr > thresh` with thresh rounded to float).
 * `order` is the score-descending permutation (computed by the caller so that the tie rule
 * is explicit); returns the number of kept indices written to keep_out (input-order ids). */
int orc_nms(const float* dets, int n, int stride, const int64_t* order, double thresh,
            int rule_ge, int32_t* keep_out, uint8_t* suppressed /* n bytes scratch */)
{
    int nkeep = 0;
    memset(suppressed, 0, (size_t)n);
    float thresh_f = (float)thresh;
    for (int _i = 0; _i < n; ++_i) {
        int i = (int)order[_i];
        if (suppressed[i]) continue;
        keep_out[nkeep++] = i;
        float ix1 = dets[i * stride + 0], iy1 = dets[i * stride + 1];
        float ix2 = dets[i * stride + 2], iy2 = dets[i * stride + 3];
        float iarea = (ix2 - ix1 + 1) * (iy2 - iy1 + 1);
        for (int _j = _i + 1; _j < n; ++_j) {
            int j = (int)order[_j];
            if (suppressed[j]) continue;
            float jx1 = dets[j * stride + 0], jy1 = dets[j * stride + 1];
            float jx2 = dets[j * stride + 2], jy2 = dets[j * stride + 3];
            float jarea = (jx2 - jx1 + 1) * (jy2 - jy1 + 1);
            float xx1 = ix1 >= jx1 ? ix1 : jx1;
            float yy1 = iy1 >= jy1 ? iy1 : jy1;
            float xx2 = ix2 <= jx2 ? ix2 : jx2;
            float yy2 = iy2 <= jy2 ? iy2 : jy2;
            float w = xx2 - xx1 + 1; if (!(w >= 0.0f)) w = 0.0f;
            float h = yy2 - yy1 + 1; if (!(h >= 0.0f)) h = 0.0f;
            float inter = w * h;
            float ovr = inter / (iarea + jarea - inter);
            if (rule_ge ? ((double)ovr >= thresh) : (ovr > thresh_f)) suppressed[j] = 1;
        }
    }
    return nkeep;
}

/* lib/utils/nms.pyx:70-123 nms_new: cpu_nms plus the containment rule ovr1/ovr2 > 0.95. */
int orc_nms_new(const float* dets, int n, int stride, const int64_t* order, double thresh,
                int32_t* keep_out, uint8_t* suppressed)
{
    int nkeep = 0;
    memset(suppressed, 0, (size_t)n);
    for (int _i = 0; _i < n; ++_i) {
        int i = (int)order[_i];
        if (suppressed[i]) continue;
        keep_out[nkeep++] = i;
        float ix1 = dets[i * stride + 0], iy1 = dets[i * stride + 1];
        float ix2 = dets[i * stride + 2], iy2 = dets[i * stride + 3];
        float iarea = (ix2 - ix1 + 1) * (iy2 - iy1 + 1);
        for (int _j = _i + 1; _j < n; ++_j) {
            int j = (int)order[_j];
            if (suppressed[j]) continue;
            float jx1 = dets[j * stride + 0], jy1 = dets[j * stride + 1];
            float jx2 = dets[j * stride + 2], jy2 = dets[j * stride + 3];
            float jarea = (jx2 - jx1 + 1) * (jy2 - jy1 + 1);
            float xx1 = ix1 >= jx1 ? ix1 : jx1;
            float yy1 = iy1 >= jy1 ? iy1 : jy1;
            float xx2 = ix2 <= jx2 ? ix2 : jx2;
            float yy2 = iy2 <= jy2 ? iy2 : jy2;
            float w = xx2 - xx1 + 1; if (!(w >= 0.0f)) w = 0.0f;
            float h = yy2 - yy1 + 1; if (!(h >= 0.0f)) h = 0.0f;
            float inter = w * h;
            float ovr = inter / (iarea + jarea - inter);
            float ovr1 = inter / iarea;
            float ovr2 = inter / jarea;
            if ((double)ovr >= thresh || (double)ovr1 > 0.95 || (double)ovr2 > 0.95) suppressed[j] = 1;
        }
    }
    return nkeep;
}

/* lib/utils/bbox.pyx:15-55 -- float64, +1 pixel convention. out is (N,K) row-major. */
void orc_bbox_overlaps(const double* boxes, int N, const double* query, int K, double* out)
{
    memset(out, 0, sizeof(double) * (size_t)N * (size_t)K);
    for (int k = 0; k < K; ++k) {
        const double* q = query + 4 * k;
        double box_area = (q[2] - q[0] + 1) * (q[3] - q[1] + 1);
        for (int n = 0; n < N; ++n) {
            const double* b = boxes + 4 * n;
            double iw = (b[2] < q[2] ? b[2] : q[2]) - (b[0] > q[0] ? b[0] : q[0]) + 1;
            if (iw > 0) {
                double ih = (b[3] < q[3] ? b[3] : q[3]) - (b[1] > q[1] ? b[1] : q[1]) + 1;
                if (ih > 0) {
                    double ua = (b[2] - b[0] + 1) * (b[3] - b[1] + 1) + box_area - iw * ih;
                    out[(size_t)n * K + k] = iw * ih / ua;
                }
            }
        }
    }
}

/* lib/roi_pooling_layer/roi_pooling_op.cc:123-182 (CPU forward), NHWC float32.
 * rois (R,5) = [batch, x1, y1, x2, y2]; top/argmax (R,PH,PW,C). */
void orc_roi_pool_fwd(const float* data, int B, int H, int W, int C, const float* rois, int R,
                      int PH, int PW, float scale, float* top, int32_t* argmax)
{
    (void)B;
    for (int n = 0; n < R; ++n) {
        const float* r = rois + 5 * n;
        int bi = (int)r[0];
        int rsw = (int)roundf(r[1] * scale), rsh = (int)roundf(r[2] * scale);
        int rew = (int)roundf(r[3] * scale), reh = (int)roundf(r[4] * scale);
        int rw = rew - rsw + 1; if (rw < 1) rw = 1;
        int rh = reh - rsh + 1; if (rh < 1) rh = 1;
        float bsh = (float)rh / (float)PH, bsw = (float)rw / (float)PW;
        const float* d = data + (size_t)bi * C * H * W;
        for (int ph = 0; ph < PH; ++ph)
            for (int pw = 0; pw < PW; ++pw) {
                int hs = (int)floorf(ph * bsh), ws = (int)floorf(pw * bsw);
                int he = (int)ceilf((ph + 1) * bsh), we = (int)ceilf((pw + 1) * bsw);
                hs += rsh; he += rsh; ws += rsw; we += rsw;
                hs = hs < 0 ? 0 : (hs > H ? H : hs); he = he < 0 ? 0 : (he > H ? H : he);
                ws = ws < 0 ? 0 : (ws > W ? W : ws); we = we < 0 ? 0 : (we > W ? W : we);
                int empty = (he <= hs) || (we <= ws);
                for (int c = 0; c < C; ++c) {
                    float mv = empty ? 0.f : -FLT_MAX;
                    int mi = -1;
                    for (int h = hs; h < he; ++h)
                        for (int w = ws; w < we; ++w) {
                            int idx = (h * W + w) * C + c;
                            if (d[idx] > mv) { mv = d[idx]; mi = idx; }
                        }
                    size_t o = (((size_t)n * PH + ph) * PW + pw) * C + c;
                    top[o] = mv; argmax[o] = mi;
                }
            }
    }
}

/* lib/roi_pooling_layer/roi_pooling_op.cc:369-444 (CPU backward): gather per input element,
 * rois visited in index order, (ph,pw) in row-major order -> a fixed float32 sum order. */
void orc_roi_pool_bwd(const float* rois, int R, const int32_t* argmax, const float* dtop, int B, int H,
                      int W, int C, int PH, int PW, float scale, float* ddata)
{
    for (int n = 0; n < B; ++n)
        for (int h = 0; h < H; ++h)
            for (int w = 0; w < W; ++w)
                for (int c = 0; c < C; ++c) {
                    float g = 0.f;
                    for (int rn = 0; rn < R; ++rn) {
                        const float* r = rois + 5 * rn;
                        if ((int)r[0] != n) continue;
                        int rsw = (int)roundf(r[1] * scale), rsh = (int)roundf(r[2] * scale);
                        int rew = (int)roundf(r[3] * scale), reh = (int)roundf(r[4] * scale);
                        if (!(w >= rsw && w <= rew && h >= rsh && h <= reh)) continue;
                        int rw = rew - rsw + 1; if (rw < 1) rw = 1;
                        int rh = reh - rsh + 1; if (rh < 1) rh = 1;
                        float bsh = (float)rh / (float)PH, bsw = (float)rw / (float)PW;
                        int phs = (int)floorf((float)(h - rsh) / bsh), phe = (int)ceilf((float)(h - rsh + 1) / bsh);
                        int pws = (int)floorf((float)(w - rsw) / bsw), pwe = (int)ceilf((float)(w - rsw + 1) / bsw);
                        phs = phs < 0 ? 0 : (phs > PH ? PH : phs); phe = phe < 0 ? 0 : (phe > PH ? PH : phe);
                        pws = pws < 0 ? 0 : (pws > PW ? PW : pws); pwe = pwe < 0 ? 0 : (pwe > PW ? PW : pwe);
                        size_t off = (size_t)rn * PH * PW * C;
                        for (int ph = phs; ph < phe; ++ph)
                            for (int pw = pws; pw < pwe; ++pw)
                                if (argmax[off + (ph * PW + pw) * C + c] == (h * W + w) * C + c)
                                    g += dtop[off + (ph * PW + pw) * C + c];
                    }
                    ddata[(((size_t)n * H + h) * W + w) * C + c] = g;
                }
}

/* tools/read_lidar.py:10-115 restated point-by-point.  For every height slice i (bounds
 * lo[i] <= z < hi[i], doubles, z widened from float -- numpy-2 promotion) in ascending
 * order and every point in file order the reference assigns top[row,col,i] = z - h0 and
 * top[row,col,zmax] = reflectance, so the last writer in (slice, file) order wins.  This
 * loop performs exactly those writes one at a time. top is (H,W,C) zero-initialised here. */
void orc_raster(const float* pts, int n, int H, int W, int C, int nslices, const double* lo, const double* hi,
                float res, float fwd0, float fwd1, float side0, float side1, float h0, int xoff, int yoff,
                float* top)
{
    memset(top, 0, sizeof(float) * (size_t)H * W * C);
    int zmax = C - 1;
    for (int i = 0; i < nslices; ++i)
        for (int p = 0; p < n; ++p) {
            float x = pts[4 * p], y = pts[4 * p + 1], z = pts[4 * p + 2], r = pts[4 * p + 3];
            if (!(x > fwd0 && x < fwd1 && y > -side1 && y < -side0)) continue;
            if (!((double)z >= lo[i] && (double)z < hi[i])) continue;
            int xi = (int)(-y / res) - xoff; /* col */
            int yi = (int)(-x / res) + yoff; /* row */
            if (yi < 0) yi += H;             /* numpy negative-index wrap (never hit for valid ranges) */
            if (xi < 0) xi += W;
            float* cell = top + ((size_t)yi * W + xi) * C;
            cell[i] = z - h0;
            cell[zmax] = r;
        }
}

/* lib/utils/transform.py:483-500 (live lidar_cnr_to_img) + :369-386 (_single).
 * corners (N,24) float32 [x0..7,y0..7,z0..7]; M = (P2 . R0) . Tr as float32 3x4 (computed by the
 * caller with numpy float32 matmuls exactly as the reference does).  The homogeneous row is
 * zeros (:381) so column 3 of M never contributes.  Products in double, divide by row 2
 * (no abs), min/max over the 8 corners, C cast to int32 (x86 cvttsd2si: NaN/inf -> INT_MIN). */
static int32_t cast_i32(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT32_MIN; /* also NaN */
    return (int32_t)v;
}
void orc_cnr_to_img(const float* corners, int n, const float* M, int32_t* out)
{
    for (int b = 0; b < n; ++b) {
        const float* c = corners + 24 * b;
        double xmin = 0, xmax = 0, ymin = 0, ymax = 0;
        int nan_x = 0, nan_y = 0;
        for (int k = 0; k < 8; ++k) {
            double X = c[k], Y = c[8 + k], Z = c[16 + k];
            double u[3];
            for (int r = 0; r < 3; ++r) {
                double acc = (double)M[4 * r] * X;
                acc = acc + (double)M[4 * r + 1] * Y;
                acc = acc + (double)M[4 * r + 2] * Z;
                acc = acc + (double)M[4 * r + 3] * 0.0;
                u[r] = acc;
            }
            double px = u[0] / u[2], py = u[1] / u[2];
            if (px != px) nan_x = 1;
            if (py != py) nan_y = 1;
            if (k == 0) { xmin = xmax = px; ymin = ymax = py; }
            else {
                if (px < xmin) xmin = px;
                if (px > xmax) xmax = px;
                if (py < ymin) ymin = py;
                if (py > ymax) ymax = py;
            }
        }
        if (nan_x) xmin = xmax = NAN; /* np.min/np.max propagate NaN */
        if (nan_y) ymin = ymax = NAN;
        out[4 * b + 0] = cast_i32(xmin);
        out[4 * b + 1] = cast_i32(ymin);
        out[4 * b + 2] = cast_i32(xmax);
        out[4 * b + 3] = cast_i32(ymax);
    }
}
