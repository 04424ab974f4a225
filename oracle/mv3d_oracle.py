"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the MV3D per-frame hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (mv3d_tf_b200) never does.

Every function restates a piece of the reference (leeyevi/MV3D_TF @ 4e1bb30) and cites the
file:line it follows (paths relative to the reference root).  Scalar loops live in
oracle/oracle_c.c (gcc); everything here is numpy with the reference's dtypes.

PARITY PIN: tests/test_oracle_vs_reference.py runs the reference's own sources (through
oracle/ref_shim.py, mechanical py2->py3 patches only) against these functions on seeded inputs
in the dev container, and tests/golden/*.npz (made by tests/golden/make_golden.py from the
*reference*, not from this file) pin them wherever the reference tree is absent.
The TensorFlow part of the path (conv / fc / softmax) cannot run here (TF 1.0, unpinned,
not installable): oracle/net_oracle.py restates it in torch-CPU fp32 -- parity UNPINNED there.

Generalisation beyond the reference: `BevGeometry` parameterises the BEV extent.  With the
default (0..60 m, -30..30 m, 0.1 m) every function is the reference's arithmetic verbatim;
other extents use the de-swapped Xn/Yn form (SURVEY.md section 8a note) which is identical on
the square reference grid.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle_c.so")
        if not os.path.exists(path):
            from . import build as _b  # type: ignore

            path = _b.build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_nms.restype = ctypes.c_int
        _LIB.orc_nms_new.restype = ctypes.c_int
    return _LIB


def _p(a, t=ctypes.c_void_p):
    return a.ctypes.data_as(t)


# ----------------------------------------------------------------------------------------
# geometry constants  (lib/utils/transform.py:3-11)
# ----------------------------------------------------------------------------------------
LIDAR_HEIGHT = 1.73
CAR_HEIGHT = 1.56


@dataclass(frozen=True)
class BevGeometry:
    """BEV extent.  Defaults = lib/utils/transform.py:3-7 (the only extent the reference has)."""

    x_min: float = 0
    x_max: float = 60
    y_min: float = -30
    y_max: float = 30
    res: float = 0.1

    @property
    def xn(self) -> int:  # transform.py:10  (note 60 // 0.1 == 599.0 -> 600)
        return int((self.x_max - self.x_min) // self.res) + 1

    @property
    def yn(self) -> int:  # transform.py:11
        return int((self.y_max - self.y_min) // self.res) + 1


REF_GEOMETRY = BevGeometry()
CFG_GEOMETRY = BevGeometry(0, 70, -40, 40, 0.1)  # BASELINE.json 700x800 grid


# ----------------------------------------------------------------------------------------
# a1. BEV raster  (tools/read_lidar.py:10-115)
# ----------------------------------------------------------------------------------------
def raster_params(res, zres, side_range, fwd_range, height_range):
    """Host-side scalars of point_cloud_2_top (read_lidar.py:49-53,80,102-103), computed with the
    same Python/numpy expressions as the reference so every rounding matches."""
    x_max = int((side_range[1] - side_range[0]) / res)
    y_max = int((fwd_range[1] - fwd_range[0]) / res)
    z_max = int((height_range[1] - height_range[0]) / zres)
    heights = np.arange(height_range[0], height_range[1], zres)  # float64 slice lower bounds
    lo = np.asarray(heights, dtype=np.float64)
    hi = np.asarray([h + zres for h in heights], dtype=np.float64)
    xoff = int(np.floor(side_range[0] / res))
    yoff = int(np.floor(fwd_range[1] / res))
    return dict(H=y_max + 1, W=x_max + 1, C=z_max + 1, nslices=len(heights), lo=lo, hi=hi, xoff=xoff, yoff=yoff)


def point_cloud_2_top(points, res=0.1, zres=0.3, side_range=(-10.0, 10.0), fwd_range=(-10.0, 10.0),
                      height_range=(-2.0, 2.0)):
    """tools/read_lidar.py:10-115 -- same signature, same (H,W,C) float32 result."""
    prm = raster_params(res, zres, side_range, fwd_range, height_range)
    pts = np.ascontiguousarray(points[:, :4], dtype=np.float32)
    top = np.empty((prm["H"], prm["W"], prm["C"]), dtype=np.float32)
    f = ctypes.c_float
    _lib().orc_raster(_p(pts), ctypes.c_int(pts.shape[0]), ctypes.c_int(prm["H"]), ctypes.c_int(prm["W"]),
                      ctypes.c_int(prm["C"]), ctypes.c_int(prm["nslices"]), _p(prm["lo"]), _p(prm["hi"]),
                      f(res), f(fwd_range[0]), f(fwd_range[1]), f(side_range[0]), f(side_range[1]),
                      f(height_range[0]), ctypes.c_int(prm["xoff"]), ctypes.c_int(prm["yoff"]), _p(top))
    return top


# ----------------------------------------------------------------------------------------
# a4. anchors  (lib/rpn_msr/generate_anchors.py:37-51, proposal_layer_tf.py:79-95)
# ----------------------------------------------------------------------------------------
def generate_anchors_bv(base_size=((3.9, 1.6), (1.0, 0.6)), res=0.1):
    """generate_anchors.py:37-51: integer (x1,y1,x2,y2) windows, each base plus its transpose."""
    rows = []
    for length, width in base_size:
        bl, bw = int(length / res), int(width / res)
        rows.append([-(bl // 2), -(bw // 2), bl - bl // 2, bw - bw // 2])
    base = np.asarray(rows, dtype=np.int64)
    return np.vstack((base, base[:, [1, 0, 3, 2]]))


def enumerate_anchors(height, width, feat_stride=8, base=None):
    """proposal_layer_tf.py:79-95 / anchor_target_layer_tf.py:76-89: (H*W*A, 4) int64, order (h, w, a)."""
    base = generate_anchors_bv() if base is None else base
    sx = np.arange(width, dtype=np.int64) * feat_stride
    sy = np.arange(height, dtype=np.int64) * feat_stride
    shifts = np.stack(np.broadcast_arrays(sx[None, :], sy[:, None], sx[None, :], sy[:, None]), axis=-1)
    return (shifts.reshape(-1, 1, 4) + base[None, :, :]).reshape(-1, 4)


# ----------------------------------------------------------------------------------------
# a5-a8. box geometry
# ----------------------------------------------------------------------------------------
def bv_anchor_to_lidar(anchors, geom: BevGeometry = REF_GEOMETRY):
    """transform.py:89-111 with _bv_to_lidar_coords :81-87.  float64 result (z, h columns hold
    float32-valued constants).  The reference multiplies bv x by Xn and bv y by Yn (swapped);
    on its square grid both are 600.  Other grids use the consistent pairing."""
    a = anchors
    length = ((a[:, 3] - a[:, 1]).reshape(-1, 1)) * geom.res
    width = ((a[:, 2] - a[:, 0]).reshape(-1, 1)) * geom.res
    cx = ((a[:, 0] + a[:, 2]) / 2.0).reshape(-1, 1)
    cy = ((a[:, 1] + a[:, 3]) / 2.0).reshape(-1, 1)
    n_for_bvx, n_for_bvy = (geom.xn, geom.yn) if geom == REF_GEOMETRY else (geom.yn, geom.xn)
    y = n_for_bvx * geom.res - (cx + 0.5) * geom.res + geom.y_min
    x = n_for_bvy * geom.res - (cy + 0.5) * geom.res + geom.x_min
    n = a.shape[0]
    h = np.ones((n, 1), dtype=np.float32) * CAR_HEIGHT
    z = np.ones((n, 1), dtype=np.float32) * -(LIDAR_HEIGHT - CAR_HEIGHT / 2.0)
    return np.hstack((x, y, z, length, width, h))


def bbox_transform_inv_3d(boxes, deltas):
    """bbox_transform.py:108-155 (single class): float32, separate mul and add roundings."""
    if boxes.shape[0] == 0:
        return np.zeros((0, deltas.shape[1]), dtype=deltas.dtype)
    b = boxes.astype(deltas.dtype, copy=False)
    out = np.zeros(deltas.shape, dtype=deltas.dtype)
    for k in range(3):  # x uses l, y uses w, z uses h  (:131-133)
        out[:, k] = deltas[:, k] * b[:, 3 + k] + b[:, k]
        out[:, 3 + k] = np.exp(deltas[:, 3 + k]) * b[:, 3 + k]
    return out


def bbox_transform_3d(ex, gt):
    """bbox_transform.py:32-58: note x is normalised by WIDTH and y by LENGTH (asymmetric to the inverse)."""
    dx = (gt[:, 0] - ex[:, 0]) / ex[:, 4]
    dy = (gt[:, 1] - ex[:, 1]) / ex[:, 3]
    dz = (gt[:, 2] - ex[:, 2]) / ex[:, 5]
    dl = np.log(gt[:, 3] / ex[:, 3])
    dw = np.log(gt[:, 4] / ex[:, 4])
    dh = np.log(gt[:, 5] / ex[:, 5])
    return np.vstack((dx, dy, dz, dl, dw, dh)).transpose()


def lidar_to_bv_coord(x, y, geom: BevGeometry = REF_GEOMETRY):
    """transform.py:13-20.  `//` here is numpy's float64 floor-division (npy_divmod)."""
    xx = geom.yn - (y - geom.y_min) // geom.res
    yy = geom.xn - (x - geom.x_min) // geom.res
    return xx, yy


def lidar_3d_to_bv(rois_3d, geom: BevGeometry = REF_GEOMETRY):
    """transform.py:113-142 (2-D branch): f32 corner sums stored to f64, `//`, cast f32."""
    r = np.zeros((rois_3d.shape[0], 4))
    r[:, 0] = rois_3d[:, 0] + rois_3d[:, 3] * 0.5
    r[:, 1] = rois_3d[:, 1] + rois_3d[:, 4] * 0.5
    r[:, 2] = rois_3d[:, 0] - rois_3d[:, 3] * 0.5
    r[:, 3] = rois_3d[:, 1] - rois_3d[:, 4] * 0.5
    x1, y1 = lidar_to_bv_coord(r[:, 0], r[:, 1], geom)
    x2, y2 = lidar_to_bv_coord(r[:, 2], r[:, 3], geom)
    return np.stack((x1, y1, x2, y2), axis=1).astype(np.float32)


_SX = np.array([1, 1, -1, -1, 1, 1, -1, -1])
_SY = np.array([1, -1, -1, 1, 1, -1, -1, 1])
_SZ = np.array([-1, -1, -1, -1, 1, 1, 1, 1])


def lidar_3d_to_corners(p):
    """transform.py:290-315: (N,24) [x0..7, y0..7, z0..7], dtype of the input (f32 on the hot path).
    +-(v/2.) then + centre, two roundings, as the reference."""
    dt = p.dtype
    hl, hw, hh = (p[:, 3:4] / 2.0), (p[:, 4:5] / 2.0), (p[:, 5:6] / 2.0)
    xs = np.where(_SX > 0, hl, -hl).astype(dt) + p[:, 0:1]
    ys = np.where(_SY > 0, hw, -hw).astype(dt) + p[:, 1:2]
    zs = np.where(_SZ > 0, hh, -hh).astype(dt) + p[:, 2:3]
    return np.hstack((xs, ys, zs)).astype(dt, copy=False)


def projection_matrix(Tr, R0, P2):
    """transform.py:371-384: mat2 = (P2 . R0[4x3]) . Tr, evaluated in the dtype the arrays arrive in
    (float32 at the py_func boundary, MV3D_test.py:18)."""
    Tr = np.asarray(Tr).reshape(3, 4)
    R0 = np.asarray(R0).reshape(4, 3)
    P2 = np.asarray(P2).reshape(3, 4)
    return np.dot(np.dot(P2, R0), Tr)


def lidar_cnr_to_img(corners, Tr, R0, P2):
    """transform.py:483-500 (the live, second definition) -> (N,4) int32 [xmin,ymin,xmax,ymax]."""
    M = np.ascontiguousarray(projection_matrix(Tr, R0, P2), dtype=np.float32)
    c = np.ascontiguousarray(corners, dtype=np.float32)
    out = np.empty((c.shape[0], 4), dtype=np.int32)
    _lib().orc_cnr_to_img(_p(c), ctypes.c_int(c.shape[0]), _p(M), _p(out))
    return out


def lidar_cnr_to_img_loop(corners, Tr, R0, P2):
    """Same function in the reference's own shape -- one box at a time, three np.dot each
    (transform.py:369-386,483-500).  Used by the timing legs so that the CPU baseline pays what
    the reference pays; tests check it equals the C loop above."""
    Tr = np.asarray(Tr).reshape(3, 4)
    R0 = np.asarray(R0).reshape(4, 3)
    P2 = np.asarray(P2).reshape(3, 4)
    out = np.zeros((corners.shape[0], 4))
    zeros8 = np.zeros(8)
    with np.errstate(all="ignore"):
        for i in range(corners.shape[0]):
            c = np.vstack((corners[i].reshape(3, 8), zeros8))
            uvw = np.dot(np.dot(np.dot(P2, R0), Tr), c)
            uvw = uvw / uvw[2]
            out[i] = (np.min(uvw[0]), np.min(uvw[1]), np.max(uvw[0]), np.max(uvw[1]))
        return out.astype(np.int32)


def clip_boxes(boxes, im_shape):
    """bbox_transform.py:178-191 (in place, like the reference)."""
    boxes[:, 0::4] = np.maximum(np.minimum(boxes[:, 0::4], im_shape[1] - 1), 0)
    boxes[:, 1::4] = np.maximum(np.minimum(boxes[:, 1::4], im_shape[0] - 1), 0)
    boxes[:, 2::4] = np.maximum(np.minimum(boxes[:, 2::4], im_shape[1] - 1), 0)
    boxes[:, 3::4] = np.maximum(np.minimum(boxes[:, 3::4], im_shape[0] - 1), 0)
    return boxes


# ----------------------------------------------------------------------------------------
# a10-a11. sort + NMS
# ----------------------------------------------------------------------------------------
def argsort_desc(scores):
    """proposal_layer_tf.py:161 / cpu_nms.pyx:25 `argsort()[::-1]` with the tie rule fixed:
    stable ascending sort then reversed => ties come out higher-index-first."""
    return np.argsort(scores.ravel(), kind="stable")[::-1]


def nms(dets, thresh, rule="ge"):
    """cpu_nms.pyx:17-68 (rule 'ge') or nms_kernel.cu:24-78 + host reduce :124-139 (rule 'gt').
    Returns kept indices into `dets`, score-descending."""
    if dets.shape[0] == 0:  # nms_wrapper.py:16-17
        return []
    d = np.ascontiguousarray(dets, dtype=np.float32)
    order = np.ascontiguousarray(argsort_desc(d[:, 4]), dtype=np.int64)
    keep = np.empty(d.shape[0], dtype=np.int32)
    scratch = np.empty(d.shape[0], dtype=np.uint8)
    n = _lib().orc_nms(_p(d), ctypes.c_int(d.shape[0]), ctypes.c_int(d.shape[1]), _p(order),
                       ctypes.c_double(float(thresh)), ctypes.c_int(1 if rule == "ge" else 0), _p(keep), _p(scratch))
    return keep[:n].tolist()


def nms_new(dets, thresh):
    """lib/utils/nms.pyx:70-123."""
    if dets.shape[0] == 0:
        return []
    d = np.ascontiguousarray(dets, dtype=np.float32)
    order = np.ascontiguousarray(argsort_desc(d[:, 4]), dtype=np.int64)
    keep = np.empty(d.shape[0], dtype=np.int32)
    scratch = np.empty(d.shape[0], dtype=np.uint8)
    n = _lib().orc_nms_new(_p(d), ctypes.c_int(d.shape[0]), ctypes.c_int(d.shape[1]), _p(order),
                           ctypes.c_double(float(thresh)), _p(keep), _p(scratch))
    return keep[:n].tolist()


def bbox_overlaps(boxes, query_boxes):
    """lib/utils/bbox.pyx:15-55: (N,K) float64 IoU with the +1 pixel convention."""
    b = np.ascontiguousarray(boxes, dtype=np.float64)
    q = np.ascontiguousarray(query_boxes, dtype=np.float64)
    out = np.empty((b.shape[0], q.shape[0]), dtype=np.float64)
    _lib().orc_bbox_overlaps(_p(b), ctypes.c_int(b.shape[0]), _p(q), ctypes.c_int(q.shape[0]), _p(out))
    return out


# ----------------------------------------------------------------------------------------
# a12. proposal_layer_3d  (lib/rpn_msr/proposal_layer_tf.py:25-202)
# ----------------------------------------------------------------------------------------
# cfg values the layer reads (config.py:138-147,187-196 defaults; yml overlay in parentheses)
RPN_CFG = {
    "TRAIN": dict(RPN_PRE_NMS_TOP_N=12000, RPN_POST_NMS_TOP_N=2000, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=5),
    "TEST": dict(RPN_PRE_NMS_TOP_N=6000, RPN_POST_NMS_TOP_N=300, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=5),
    "TEST_DEFAULT": dict(RPN_PRE_NMS_TOP_N=12000, RPN_POST_NMS_TOP_N=2000, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=5),
}


def proposal_stages(rpn_cls_prob_reshape, rpn_bbox_pred, im_info, calib, feat_stride=8,
                    geom: BevGeometry = REF_GEOMETRY, img_size=(375, 1242), project="c"):
    """Steps 1-3 of proposal_layer_3d (:63-151) for ALL anchors, no compaction: returns a dict with
    per-anchor scores, proposals_3d (f32), proposals_bv (clipped, f32), proposals_img (int32) and the
    two keep masks.  Split out so the CUDA decode kernel can be compared stage by stage."""
    A = generate_anchors_bv().shape[0]
    height, width = rpn_cls_prob_reshape.shape[1:3]
    scores = rpn_cls_prob_reshape.reshape(1, height, width, A, 2)[..., 1].reshape(-1)
    deltas = rpn_bbox_pred.reshape(-1, 6)
    anchors = enumerate_anchors(height, width, feat_stride)
    anchors_3d = bv_anchor_to_lidar(anchors, geom)
    with np.errstate(all="ignore"):
        p3d = bbox_transform_inv_3d(anchors_3d, deltas)
        pbv = lidar_3d_to_bv(p3d, geom)
        cnr = lidar_3d_to_corners(p3d)
        fn = lidar_cnr_to_img if project == "c" else lidar_cnr_to_img_loop
        pimg = fn(cnr, calib[3], calib[2], calib[0])
        info = np.asarray(im_info).reshape(-1, 3)[0]
        pbv = clip_boxes(pbv, info[:2])
        min_size = 5 * info[2]
        ws = pbv[:, 2] - pbv[:, 0] + 1
        hs = pbv[:, 3] - pbv[:, 1] + 1
        keep_size = (ws >= min_size) & (hs >= min_size)  # _filter_boxes :336-341
        pad = 50  # _filter_img_boxes :343-352, image size hard-coded at the call site :147
        keep_img = ((-pad <= pimg[:, 0]) & (pimg[:, 2] <= img_size[1] + pad) & (-pad <= pimg[:, 1]) &
                    (pimg[:, 3] <= img_size[0] + pad))
    return dict(scores=scores, anchors=anchors, anchors_3d=anchors_3d.astype(np.float32), p3d=p3d, pbv=pbv,
                pimg=pimg, keep_size=keep_size, keep_img=keep_img)


def proposal_layer_3d(rpn_cls_prob_reshape, rpn_bbox_pred, im_info, calib, cfg_key, _feat_stride=(8,),
                      anchor_scales=(1.0, 1.0), cfg=None, geom: BevGeometry = REF_GEOMETRY, img_size=(375, 1242),
                      nms_rule="ge", project="c", return_stages=False):
    """proposal_layer_tf.py:25-202 -> (blob_bv (R,5), blob_img (R,5), blob_3d (R,7)) float32."""
    assert rpn_cls_prob_reshape.shape[0] == 1, "Only single item batches are supported"
    c = (cfg or RPN_CFG)[cfg_key]
    st = proposal_stages(rpn_cls_prob_reshape, rpn_bbox_pred, im_info, calib, int(_feat_stride[0]), geom, img_size,
                         project)
    keep = np.where(st["keep_size"] & st["keep_img"])[0]  # two successive filters == one AND, order kept
    pbv, p3d, pimg, scores = st["pbv"][keep], st["p3d"][keep], st["pimg"][keep], st["scores"][keep]
    order = argsort_desc(scores)
    if c["RPN_PRE_NMS_TOP_N"] > 0:
        order = order[: c["RPN_PRE_NMS_TOP_N"]]
    pbv, p3d, pimg, scores = pbv[order], p3d[order], pimg[order], scores[order]
    k = nms(np.hstack((pbv, scores.reshape(-1, 1))), c["RPN_NMS_THRESH"], nms_rule)
    if c["RPN_POST_NMS_TOP_N"] > 0:
        k = k[: c["RPN_POST_NMS_TOP_N"]]
    k = np.asarray(k, dtype=np.int64)
    pbv, p3d, pimg, scores = pbv[k], p3d[k], pimg[k], scores[k]
    z = np.zeros((pbv.shape[0], 1), dtype=np.float32)
    blobs = (np.hstack((z, pbv.astype(np.float32))), np.hstack((z, pimg.astype(np.float32))),
             np.hstack((z, p3d.astype(np.float32))))
    if return_stages:
        st.update(keep=keep, order=order, nms_keep=k, final_scores=scores, anchor_index=keep[order][k])
        return blobs, st
    return blobs


# ----------------------------------------------------------------------------------------
# a15-a16. ROI pooling  (lib/roi_pooling_layer/roi_pooling_op.cc:123-182, :369-444)
# ----------------------------------------------------------------------------------------
def roi_pool_fwd(data, rois, pooled_h=7, pooled_w=7, spatial_scale=0.125):
    d = np.ascontiguousarray(data, dtype=np.float32)
    r = np.ascontiguousarray(rois, dtype=np.float32)
    B, H, W, C = d.shape
    top = np.empty((r.shape[0], pooled_h, pooled_w, C), dtype=np.float32)
    arg = np.empty(top.shape, dtype=np.int32)
    i = ctypes.c_int
    _lib().orc_roi_pool_fwd(_p(d), i(B), i(H), i(W), i(C), _p(r), i(r.shape[0]), i(pooled_h), i(pooled_w),
                            ctypes.c_float(spatial_scale), _p(top), _p(arg))
    return top, arg


def roi_pool_bwd(data_shape, rois, argmax, dtop, pooled_h=7, pooled_w=7, spatial_scale=0.125):
    r = np.ascontiguousarray(rois, dtype=np.float32)
    a = np.ascontiguousarray(argmax, dtype=np.int32)
    g = np.ascontiguousarray(dtop, dtype=np.float32)
    B, H, W, C = data_shape
    out = np.empty(data_shape, dtype=np.float32)
    i = ctypes.c_int
    _lib().orc_roi_pool_bwd(_p(r), i(r.shape[0]), _p(a), _p(g), i(B), i(H), i(W), i(C), i(pooled_h), i(pooled_w),
                            ctypes.c_float(spatial_scale), _p(out))
    return out


# ----------------------------------------------------------------------------------------
# a13-a14. training target layers
# ----------------------------------------------------------------------------------------
TRAIN_CFG = dict(RPN_POSITIVE_OVERLAP=0.7, RPN_NEGATIVE_OVERLAP=0.5, RPN_CLOBBER_POSITIVES=False,
                 RPN_FG_FRACTION=0.25, RPN_BATCHSIZE=128, BATCH_SIZE=128, FG_FRACTION=0.25, FG_THRESH=0.7,
                 BG_THRESH_HI=0.5, BG_THRESH_LO=0.0)   # config.py defaults + faster_rcnn_end2end.yml overlay


def anchor_target_layer(rpn_cls_score, gt_boxes, gt_boxes_3d, im_info, _feat_stride=(8,), anchor_scales=(1.0, 1.0),
                        cfg=None, geom: BevGeometry = REF_GEOMETRY, rng=None, return_stages=False):
    """anchor_target_layer_tf.py:21-250.  `rng`: the numpy RandomState-like module/object whose `.choice` is drawn
    from (the reference uses the global numpy.random).  Returns (labels (N,), bbox_targets (N,6), anchors (<=128,5),
    anchors_3d (<=128,7)) float32; with return_stages also the pre-sampling labels / max_overlaps / argmax."""
    c = dict(TRAIN_CFG)
    if cfg:
        c.update(cfg)
    npr = np.random if rng is None else rng
    im_info = np.asarray(im_info).reshape(-1, 3)[0]
    assert rpn_cls_score.shape[0] == 1
    height, width = rpn_cls_score.shape[1:3]
    all_anchors = enumerate_anchors(height, width, int(np.ravel(_feat_stride)[0]))           # :76-89
    total = all_anchors.shape[0]
    inds_inside = np.where((all_anchors[:, 0] >= 0) & (all_anchors[:, 1] >= 0) &
                           (all_anchors[:, 2] < im_info[1]) & (all_anchors[:, 3] < im_info[0]))[0]   # :93-98
    anchors = all_anchors[inds_inside, :]
    labels = np.empty((len(inds_inside),), dtype=np.float32)
    labels.fill(-1)
    overlaps = bbox_overlaps(np.ascontiguousarray(anchors, dtype=np.float64),
                             np.ascontiguousarray(gt_boxes[:, :4], dtype=np.float64))        # :113-115 (cols 0-3 are read)
    argmax_overlaps = overlaps.argmax(axis=1)
    max_overlaps = overlaps[np.arange(len(inds_inside)), argmax_overlaps]
    gt_argmax_overlaps = overlaps.argmax(axis=0)
    gt_max_overlaps = overlaps[gt_argmax_overlaps, np.arange(overlaps.shape[1])]
    gt_argmax_overlaps = np.where(overlaps == gt_max_overlaps)[0]                            # :123
    if not c["RPN_CLOBBER_POSITIVES"]:
        labels[np.logical_and(0 < max_overlaps, max_overlaps < c["RPN_NEGATIVE_OVERLAP"])] = 0   # :129-130
    labels[gt_argmax_overlaps] = 1                                                           # :133
    labels[max_overlaps >= c["RPN_POSITIVE_OVERLAP"]] = 1                                    # :139
    if c["RPN_CLOBBER_POSITIVES"]:
        labels[max_overlaps < c["RPN_NEGATIVE_OVERLAP"]] = 0
    pre_labels = labels.copy()
    num_fg = int(c["RPN_FG_FRACTION"] * c["RPN_BATCHSIZE"])                                  # :146-151
    fg_inds = np.where(labels == 1)[0]
    if len(fg_inds) > num_fg:
        labels[npr.choice(fg_inds, size=(len(fg_inds) - num_fg), replace=False)] = -1
    num_bg = c["RPN_BATCHSIZE"] - np.sum(labels == 1)                                        # :154-159
    bg_inds = np.where(labels == 0)[0]
    if len(bg_inds) > num_bg:
        labels[npr.choice(bg_inds, size=(len(bg_inds) - num_bg), replace=False)] = -1
    anchors_3d = bv_anchor_to_lidar(anchors, geom)                                           # :164
    bbox_targets = bbox_transform_3d(anchors_3d, gt_boxes_3d[argmax_overlaps, :][:, :6]).astype(np.float32, copy=False)
    all_inds = np.where(labels != -1)                                                        # :169-174
    zeros = np.zeros((labels[all_inds].shape[0], 1), dtype=np.float32)
    out_anchors = np.hstack((zeros, anchors[all_inds])).astype(np.float32)
    out_anchors_3d = np.hstack((zeros, anchors_3d[all_inds])).astype(np.float32)
    labels[max_overlaps < c["RPN_NEGATIVE_OVERLAP"]] = 0                                     # :176
    num_bg = c["RPN_BATCHSIZE"] - np.sum(labels == 1)
    bg_inds = np.where(labels == 0)[0]
    if len(bg_inds) > num_bg:
        labels[npr.choice(bg_inds, size=(len(bg_inds) - num_bg), replace=False)] = -1

    def unmap(data, fill):                                                                   # :254-265
        ret = np.empty((total,) + data.shape[1:], dtype=np.float32)
        ret.fill(fill)
        ret[inds_inside] = data
        return ret
    out = (unmap(labels, -1), unmap(bbox_targets, 0), out_anchors, out_anchors_3d)
    if return_stages:
        return out + (dict(inds_inside=inds_inside, pre_labels=unmap(pre_labels, -1), max_overlaps=unmap(max_overlaps, -1),
                           argmax=argmax_overlaps),)
    return out


def bbox_transform_cnr(ex_cnr, gt_cnr):
    """bbox_transform.py:61-72."""
    diag = np.linalg.norm(gt_cnr[:, 0::8] - gt_cnr[:, 6::8], axis=1)
    return np.divide(gt_cnr - ex_cnr, diag.reshape((-1, 1)))


def proposal_target_layer_3d(rpn_rois_bv, rpn_rois_3d, gt_boxes_bv, gt_boxes_3d, gt_boxes_corners, calib, _num_classes,
                             cfg=None, rng=None, return_stages=False):
    """proposal_target_layer_tf.py:19-94 with _sample_rois_3d :227-298 -> (rois_bv (K,5), rois_img (K,5),
    labels (K,1) int32, bbox_targets (K,24*nc), rois_3d (K,7))."""
    c = dict(TRAIN_CFG)
    if cfg:
        c.update(cfg)
    npr = np.random if rng is None else rng
    zeros = np.zeros((gt_boxes_bv.shape[0], 1), dtype=gt_boxes_bv.dtype)
    all_rois = np.vstack((rpn_rois_bv, np.hstack((zeros, gt_boxes_bv[:, :-1]))))             # :38-41
    all_rois_3d = np.vstack((rpn_rois_3d, np.hstack((zeros, gt_boxes_3d[:, :-1]))))          # :42-44
    assert np.all(all_rois[:, 0] == 0)
    rois_per_image = c["BATCH_SIZE"] // 1
    fg_rois_per_image = np.round(c["FG_FRACTION"] * rois_per_image)
    overlaps = bbox_overlaps(np.ascontiguousarray(all_rois[:, 1:5], dtype=np.float64),
                             np.ascontiguousarray(gt_boxes_bv[:, :4], dtype=np.float64))     # :232-234
    gt_assignment = overlaps.argmax(axis=1)
    max_overlaps = overlaps.max(axis=1)
    labels = gt_boxes_bv[gt_assignment, 4]
    fg_inds = np.where(max_overlaps >= c["FG_THRESH"])[0]                                    # :244
    fg_n = int(min(fg_rois_per_image, fg_inds.size))
    if fg_inds.size > 0:
        fg_inds = npr.choice(fg_inds, size=fg_n, replace=False)
    bg_inds = np.where((max_overlaps < c["BG_THRESH_HI"]) & (max_overlaps >= c["BG_THRESH_LO"]))[0]   # :258-259
    bg_n = min(rois_per_image - fg_n, bg_inds.size)
    if bg_inds.size > 0:
        bg_inds = npr.choice(bg_inds, size=bg_n, replace=False)
    keep_inds = np.append(fg_inds, bg_inds).astype(np.int64)                                 # :272
    labels = labels[keep_inds]
    labels[fg_n:] = 0
    rois_bv = all_rois[keep_inds]
    rois_3d = all_rois_3d[keep_inds]
    rois_cnr = lidar_3d_to_corners(rois_3d[:, 1:7])                                          # :283
    targets = bbox_transform_cnr(rois_cnr, gt_boxes_corners[gt_assignment[keep_inds], :24])  # :293-294
    data = np.hstack((labels[:, np.newaxis], targets)).astype(np.float32, copy=False)
    clss = np.array(data[:, 0], dtype=np.uint16, copy=True)                                  # :184-192
    bbox_targets = np.zeros((clss.size, 24 * _num_classes), dtype=np.float32)
    for ind in np.where(clss > 0)[0]:
        bbox_targets[ind, 24 * clss[ind]:24 * clss[ind] + 24] = data[ind, 1:]
    calib = np.asarray(calib)
    rois_img = lidar_cnr_to_img(rois_cnr, calib[3], calib[2], calib[0])                      # :77-79
    rois_img = np.hstack((rois_bv[:, 0].reshape(-1, 1), rois_img))
    out = (rois_bv.reshape(-1, 5).astype(np.float32), rois_img.reshape(-1, 5).astype(np.float32),
           labels.reshape(-1, 1).astype(np.int32), bbox_targets.reshape(-1, _num_classes * 24).astype(np.float32),
           rois_3d.reshape(-1, 7).astype(np.float32))
    if return_stages:
        return out + (dict(keep_inds=keep_inds, fg_n=fg_n, max_overlaps=max_overlaps, gt_assignment=gt_assignment),)
    return out


def synth_gt(n_gt=6, seed=1234, geom: BevGeometry = REF_GEOMETRY):
    """SURVEY 8d synthetic ground truth: cars inside the BEV extent -> (gt_boxes_bv (G,5), gt_boxes_3d (G,7),
    gt_boxes_corners (G,25)) float32, class 1."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(geom.x_min + 5, geom.x_max - 5, n_gt)
    y = rng.uniform(geom.y_min + 5, geom.y_max - 5, n_gt)
    z = np.full(n_gt, -0.95)
    lwh = np.array([3.9, 1.6, 1.56]) * rng.uniform(0.9, 1.1, (n_gt, 3))
    if n_gt > 1:  # half of the cars are rotated by 90 degrees (length along y) so both anchor shapes get positives
        lwh[::2, [0, 1]] = lwh[::2, [1, 0]]
    p3d = np.column_stack((x, y, z, lwh)).astype(np.float32)
    bv = lidar_3d_to_bv(p3d, geom)
    cnr = lidar_3d_to_corners(p3d)
    ones = np.ones((n_gt, 1), np.float32)
    return np.hstack((bv, ones)).astype(np.float32), np.hstack((p3d, ones)).astype(np.float32), \
        np.hstack((cnr, ones)).astype(np.float32)


# ----------------------------------------------------------------------------------------
# f1. test_net post-processing (lib/fast_rcnn/test_mv.py:241-264, 420-444, 492-501)
# ----------------------------------------------------------------------------------------
def corners_to_bv(corners, geom: BevGeometry = REF_GEOMETRY):
    """transform.py:342-366 (per class block of 24)."""
    n_cls = corners.shape[1] // 24
    bv = np.zeros((corners.shape[0], 4 * n_cls))
    for i in range(n_cls):
        c = corners[:, i * 24:(i + 1) * 24]
        xmin, xmax = np.min(c[:, :8], axis=1).reshape(-1, 1), np.max(c[:, :8], axis=1).reshape(-1, 1)
        ymin, ymax = np.min(c[:, 8:16], axis=1).reshape(-1, 1), np.max(c[:, 8:16], axis=1).reshape(-1, 1)
        p = np.hstack([xmax, ymax, xmin, ymin])
        p[:, 0], p[:, 1] = lidar_to_bv_coord(p[:, 0], p[:, 1], geom)
        p[:, 2], p[:, 3] = lidar_to_bv_coord(p[:, 2], p[:, 3], geom)
        bv[:, i * 4:(i + 1) * 4] = p
    return bv


def bbox_transform_inv_cnr(boxes, deltas):
    """bbox_transform.py:157-176."""
    if boxes.shape[0] == 0:
        return np.zeros((0, deltas.shape[1]), dtype=deltas.dtype)
    boxes = boxes.astype(deltas.dtype, copy=False)
    diag = np.linalg.norm(boxes[:, 0::8] - boxes[:, 6::8], axis=1)
    deltas = np.multiply(deltas, diag.reshape((-1, 1)))
    pred = np.zeros(deltas.shape, dtype=deltas.dtype)
    for i in range(deltas.shape[1] // 24):
        pred[:, i * 24:i * 24 + 24] = deltas[:, i * 24:i * 24 + 24] + boxes
    return pred


def collect_detections(scores, boxes_bv, boxes_cnr, num_classes, thresh=0.05, nms_thresh=0.1, max_per_image=300):
    """test_mv.py:420-444 + 492-501 for one frame."""
    dets, dets_cnr = {}, {}
    for j in range(1, num_classes):
        inds = np.where(scores[:, j] > thresh)[0]
        cs = scores[inds, j]
        d = np.hstack((boxes_bv[inds, j * 4:(j + 1) * 4], cs[:, np.newaxis])).astype(np.float32, copy=False)
        dc = np.hstack((boxes_cnr[inds, j * 24:(j + 1) * 24], cs[:, np.newaxis])).astype(np.float32, copy=False)
        keep = nms(d, nms_thresh)
        dets[j], dets_cnr[j] = d[keep, :], dc[keep, :]
    if max_per_image > 0 and num_classes > 1:
        sc = np.hstack([dets[j][:, -1] for j in range(1, num_classes)])
        if len(sc) > max_per_image:
            t = np.sort(sc)[-max_per_image]
            for j in range(1, num_classes):
                k = np.where(dets[j][:, -1] >= t)[0]
                dets[j], dets_cnr[j] = dets[j][k, :], dets_cnr[j][k, :]
    return dets, dets_cnr


# ----------------------------------------------------------------------------------------
# Front view (FV).  NOT in the reference: `proposal_transform(target='fv')` returns None (network.py:313-315) and there
# is no FV raster / trunk / ROI anywhere in its tree.  BASELINE.json's north star asks for the paper's third view, so
# the functions below ARE the specification (SURVEY 8f "Front view"), written down once here and implemented in
# csrc/front_view.cu; parity for the FV branch is against this file only.
# ----------------------------------------------------------------------------------------
@dataclass(frozen=True)
class FvGeometry:
    """Cylindrical front-view map: H rows over elevation [phi_min, phi_max] (row 0 = top), W columns over azimuth
    [theta_min, theta_max) (column 0 = theta_min), angles in degrees; 64 beams x 512 columns over a 90 degree fan."""
    H: int = 64
    W: int = 512
    theta_min: float = -45.0
    theta_max: float = 45.0
    phi_min: float = -24.9
    phi_max: float = 2.0

    @property
    def dtheta(self) -> float:
        return np.deg2rad(self.theta_max - self.theta_min) / self.W

    @property
    def dphi(self) -> float:
        return np.deg2rad(self.phi_max - self.phi_min) / self.H


FV_GEOMETRY = FvGeometry()


def fv_coords(x, y, z, g: FvGeometry = FV_GEOMETRY):
    """float64 (col, row) of LiDAR points in the FV map (unfloored)."""
    x, y, z = (np.asarray(a, dtype=np.float64) for a in (x, y, z))
    col = (np.arctan2(y, x) - np.deg2rad(g.theta_min)) / g.dtheta
    row = (np.deg2rad(g.phi_max) - np.arctan2(z, np.sqrt(x * x + y * y))) / g.dphi
    return col, row


def point_cloud_2_front(points, g: FvGeometry = FV_GEOMETRY):
    """(N,>=4) float32 [x,y,z,r] -> (H,W,3) float32 [height z, distance sqrt(x^2+y^2+z^2), reflectance]; cell =
    floor of fv_coords (float64); points with x <= 0 or outside the map are dropped; LAST point in file order wins,
    as in the BEV rasteriser (tools/read_lidar.py:106-113 semantics)."""
    p = np.asarray(points, dtype=np.float32)
    out = np.zeros((g.H, g.W, 3), dtype=np.float32)
    if p.shape[0] == 0:
        return out
    x, y, z = p[:, 0].astype(np.float64), p[:, 1].astype(np.float64), p[:, 2].astype(np.float64)
    col, row = fv_coords(x, y, z, g)
    c, r = np.floor(col), np.floor(row)
    ok = (x > 0) & (c >= 0) & (c < g.W) & (r >= 0) & (r < g.H)
    ci, ri = c[ok].astype(np.int64), r[ok].astype(np.int64)
    q = p[ok]
    dist = np.sqrt(x[ok] * x[ok] + y[ok] * y[ok] + z[ok] * z[ok]).astype(np.float32)
    out[ri, ci, 0] = q[:, 2]          # numpy fancy assignment: the last duplicate wins
    out[ri, ci, 1] = dist
    out[ri, ci, 2] = q[:, 3]
    return out


def lidar_3d_to_fv(rois_3d, g: FvGeometry = FV_GEOMETRY):
    """(N,6) [x,y,z,l,w,h] float32 -> (N,4) float32 [col_min,row_min,col_max,row_max]: floor of the FV coordinates of
    the 8 corners (lidar_3d_to_corners), min / max, clamped to the map.  Corners behind the sensor (x <= 0) keep their
    atan2 value; NaN inputs give a zero box."""
    c = lidar_3d_to_corners(np.asarray(rois_3d, dtype=np.float32)).astype(np.float64)
    with np.errstate(all="ignore"):
        col, row = fv_coords(c[:, 0:8], c[:, 8:16], c[:, 16:24], g)
        col, row = np.floor(col), np.floor(row)
        bad = ~(np.isfinite(col).all(axis=1) & np.isfinite(row).all(axis=1))
        out = np.stack((np.clip(col.min(axis=1), 0, g.W - 1), np.clip(row.min(axis=1), 0, g.H - 1),
                        np.clip(col.max(axis=1), 0, g.W - 1), np.clip(row.max(axis=1), 0, g.H - 1)), axis=1)
    out[bad] = 0
    return out.astype(np.float32)


# ----------------------------------------------------------------------------------------
# synthetic inputs -- shared by tests, golden generator and bench
# ----------------------------------------------------------------------------------------
# KITTI 000008-like calibration, rows P2, P3, R0_rect (9 values, zero padded), Tr_velo_to_cam
# (layout lib/datasets/kitti_mv3d.py:63-75).  Plain KITTI numbers, not reference code.
KITTI_CALIB = np.array([
    [7.215377e+02, 0.0, 6.095593e+02, 4.485728e+01, 0.0, 7.215377e+02, 1.728540e+02, 2.163791e-01,
     0.0, 0.0, 1.0, 2.745884e-03],
    [7.215377e+02, 0.0, 6.095593e+02, -3.395242e+02, 0.0, 7.215377e+02, 1.728540e+02, 2.199936e+00,
     0.0, 0.0, 1.0, 2.729905e-03],
    [9.999239e-01, 9.837760e-03, -7.445048e-03, -9.869795e-03, 9.999421e-01, -4.278459e-03,
     7.402527e-03, 4.351614e-03, 9.999631e-01, 0.0, 0.0, 0.0],
    [7.533745e-03, -9.999714e-01, -6.166020e-04, -4.069766e-03, 1.480249e-02, 7.280733e-04,
     -9.998902e-01, -7.631618e-02, 9.998621e-01, 7.523790e-03, 1.480755e-02, -2.717806e-01],
], dtype=np.float32)


def synth_points(n, seed=1234):
    """LiDAR frame: x~U(-5,75), y~U(-45,45), z~U(-3,2), r~U(0,1), float32 (N,4)."""
    rng = np.random.default_rng(seed)
    pts = np.empty((n, 4), dtype=np.float32)
    pts[:, 0] = rng.uniform(-5, 75, n)
    pts[:, 1] = rng.uniform(-45, 45, n)
    pts[:, 2] = rng.uniform(-3, 2, n)
    pts[:, 3] = rng.uniform(0, 1, n)
    return pts


def synth_rpn_outputs(hf, wf, seed=1234, A=4):
    """Tie-free fg/bg probabilities (1,Hf,Wf,2A) and deltas (1,Hf,Wf,6A) ~ N(0,0.1), float32."""
    rng = np.random.default_rng(seed)
    n = hf * wf * A
    # fg probabilities are n DISTINCT float32 values (a permutation of (k+0.5)/n, squared to skew
    # towards background like a real RPN) so that no sort ever has to break a tie
    fg = (((rng.permutation(n) + 0.5) / n) ** 2).astype(np.float32)
    assert np.unique(fg).size == n
    prob = np.stack((np.float32(1) - fg, fg), axis=1).reshape(1, hf, wf, 2 * A)
    deltas = rng.normal(0, 0.1, (1, hf, wf, 6 * A)).astype(np.float32)
    return prob, deltas
