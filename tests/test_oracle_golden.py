"""The CPU oracle against the committed golden vectors (made from the REFERENCE's own sources by
tests/golden/make_golden.py).  Bit-exact everywhere except where np.exp (CPU-dispatch dependent,
SURVEY A3) feeds a float: those are <= 2 ulp."""
import os

import numpy as np


def _ulp_diff(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


def test_raster_golden(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "raster.npz"))
    a = oracle.point_cloud_2_top(g["points"], res=0.1, zres=0.3, side_range=(-8., 8.), fwd_range=(0., 16.),
                                 height_range=(-2, 0.4))
    b = oracle.point_cloud_2_top(g["points"], res=0.1, zres=0.1, side_range=(-8., 8.), fwd_range=(0., 12.),
                                 height_range=(-2.0, 1.5))
    assert a.shape == g["top_a"].shape and np.array_equal(a, g["top_a"])
    assert b.shape == g["top_b"].shape and np.array_equal(b, g["top_b"])


def test_raster_empty_and_outside(oracle):
    kw = dict(res=0.1, zres=0.3, side_range=(-8., 8.), fwd_range=(0., 16.), height_range=(-2, 0.4))
    top = oracle.point_cloud_2_top(np.zeros((0, 4), np.float32), **kw)
    assert top.shape == (161, 161, 9) and not top.any()
    pts = np.array([[-1, 0, 0, 1], [17, 0, 0, 1], [5, 9, 0, 1], [5, 0, 5, 1], [5, 0, -3, 1]], np.float32)
    assert not oracle.point_cloud_2_top(pts, **kw).any()


def test_anchors_golden(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "proposal.npz"))
    assert np.array_equal(oracle.generate_anchors_bv(), g["base_anchors"])
    assert oracle.generate_anchors_bv().tolist() == [[-19, -8, 20, 8], [-5, -2, 5, 3], [-8, -19, 8, 20], [-2, -5, 3, 5]]
    a = oracle.enumerate_anchors(3, 2, 8)
    assert a.shape == (24, 4) and a[4].tolist() == [-19 + 8, -8, 20 + 8, 8] and a[8].tolist() == [-19, 0, 20, 16]


def test_proposal_stages_golden(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "proposal.npz"))
    st = oracle.proposal_stages(g["prob"], g["deltas"], g["im_info"], g["calib"])
    assert np.array_equal(st["anchors_3d"], g["anchors_3d"].astype(np.float32))
    fin = np.isfinite(g["p3d"])
    assert np.array_equal(fin, np.isfinite(st["p3d"]))
    assert _ulp_diff(st["p3d"][fin], g["p3d"][fin]).max() <= 2
    # integer-valued outputs: exact wherever the float inputs agreed bit for bit
    same = (_ulp_diff(np.nan_to_num(st["p3d"]), np.nan_to_num(g["p3d"])).max(axis=1) == 0)
    assert same.mean() > 0.95
    pbv_ref = oracle.clip_boxes(g["pbv"].copy(), g["im_info"][0, :2])
    assert np.array_equal(st["pbv"][same], pbv_ref[same], equal_nan=True)
    assert np.array_equal(st["pimg"][same], g["pimg"][same])
    c = oracle.lidar_3d_to_corners(g["p3d"])
    assert np.array_equal(c, g["corners"], equal_nan=True)


def test_proposal_layer_golden(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "proposal.npz"))
    for key, pre, post in (("TEST", g["cfg"][0], g["cfg"][1]), ("TRAIN", g["cfg"][2], g["cfg"][3])):
        cfg = {key: dict(RPN_PRE_NMS_TOP_N=int(pre), RPN_POST_NMS_TOP_N=int(post), RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=5)}
        bv, img, p3d = oracle.proposal_layer_3d(g["prob"], g["deltas"], g["im_info"], g["calib"], key, cfg=cfg)
        k = key.lower()
        assert np.array_equal(bv, g[k + "_bv"])
        assert np.array_equal(img, g[k + "_img"])
        assert bv.shape[0] == p3d.shape[0] and _ulp_diff(p3d, g[k + "_3d"]).max() <= 2


def test_projection_loop_equals_c(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "proposal.npz"))
    c = g["corners"][:600]
    a = oracle.lidar_cnr_to_img(c, g["calib"][3], g["calib"][2], g["calib"][0])
    b = oracle.lidar_cnr_to_img_loop(c, g["calib"][3], g["calib"][2], g["calib"][0])
    assert np.array_equal(a, b) and np.array_equal(a, g["pimg"][:600])


def test_nms_iou_golden(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "nms_iou.npz"))
    assert oracle.nms(g["dets"], 0.7) == g["keep07"].tolist()
    assert oracle.nms(g["dets"], 0.5) == g["keep05"].tolist()
    assert oracle.nms(g["dets"], 0.1) == g["keep01"].tolist()
    assert oracle.nms_new(g["dets"], 0.3) == g["keep_new"].tolist()
    assert oracle.nms(np.zeros((0, 5), np.float32), 0.7) == []
    assert np.array_equal(oracle.bbox_overlaps(g["boxes"], g["query"]), g["iou"])
    # '>' rule (nms_kernel.cu:71) differs from '>=' only at exact-threshold IoU
    d = np.array([[0, 0, 9, 9, .9], [0, 0, 9, 4, .8]], np.float32)  # IoU = 0.5 exactly
    assert oracle.nms(d, 0.5, "ge") == [0] and oracle.nms(d, 0.5, "gt") == [0, 1]


def test_sort_tie_rule(oracle):
    s = np.array([0.5, 0.9, 0.5, 0.9, 0.1], np.float32)
    assert oracle.argsort_desc(s).tolist() == [3, 1, 2, 0, 4]


def test_npy_floor_divide_cases(oracle):
    # numpy float // is not floor(a/b)
    x = np.array([20.0, 1.0, 0.3, 60.0])
    assert (x // 0.1).tolist() == [199.0, 9.0, 2.0, 599.0]
    assert oracle.REF_GEOMETRY.xn == 600 and oracle.REF_GEOMETRY.yn == 600
    assert oracle.CFG_GEOMETRY.xn == 700 and oracle.CFG_GEOMETRY.yn == 800


def test_roi_pool_oracle_small(oracle):
    rng = np.random.default_rng(5)
    data = rng.normal(size=(1, 9, 11, 4)).astype(np.float32)
    rois = np.array([[0, 0, 0, 80, 64], [0, 16, 8, 40, 40], [0, -50, -50, -20, -20], [0, 200, 200, 300, 300]], np.float32)
    top, arg = oracle.roi_pool_fwd(data, rois)
    assert top.shape == (4, 7, 7, 4)
    assert (arg[2] == -1).all() and (top[2] == 0).all() and (arg[3] == -1).all()
    m = arg >= 0
    assert np.array_equal(top[m], data.reshape(-1)[arg[m]])
    assert top[0].max() == data.max()
    g = rng.normal(size=top.shape).astype(np.float32)
    dd = oracle.roi_pool_bwd(data.shape, rois, arg, g)
    ref = np.zeros(data.size, np.float64)
    np.add.at(ref, arg[m], g[m].astype(np.float64))
    assert np.allclose(dd.reshape(-1), ref, rtol=1e-5, atol=1e-6)


def test_target_layers_golden(oracle, golden_dir):
    """anchor_target_layer / proposal_target_layer_3d restatements against vectors made by the reference itself
    (tests/golden/make_golden_targets.py), consuming numpy's global RandomState exactly as the reference does."""
    g = np.load(os.path.join(golden_dir, "targets.npz"))
    hf, wf = int(g["hf"]), int(g["wf"])
    cls = np.zeros((1, hf, wf, 8), np.float32)
    np.random.seed(int(g["seed"]))
    labels, targets, anchors, anchors_3d = oracle.anchor_target_layer(cls, g["gt_bv"], g["gt_3d"], g["im_info"])
    assert np.array_equal(labels, g["at_labels"])
    assert np.array_equal(targets, g["at_targets"])
    assert np.array_equal(anchors, g["at_anchors"]) and np.array_equal(anchors_3d, g["at_anchors_3d"])
    np.random.seed(int(g["seed"]) + 1)
    out = oracle.proposal_target_layer_3d(g["rois_bv"], g["rois_3d"], g["gt_bv"], g["gt_3d"], g["gt_cnr"], g["calib"], 2)
    for got, key in zip(out, ("pt_rois_bv", "pt_rois_img", "pt_labels", "pt_targets", "pt_rois_3d")):
        assert got.dtype == g[key].dtype and np.array_equal(got, g[key]), key


def test_roi_pool_golden(oracle, golden_dir):
    """tests/golden/roi_pool.npz was produced by the reference's own RoiPool / RoiPoolGrad CPU kernels
    (tests/golden/make_golden_roi_pool.py); the C restatement must reproduce it bit for bit."""
    g = np.load(os.path.join(golden_dir, "roi_pool.npz"))
    top, arg = oracle.roi_pool_fwd(g["data"], g["rois"])
    assert np.array_equal(top, g["top"]) and np.array_equal(arg, g["argmax"])
    assert np.array_equal(oracle.roi_pool_bwd(g["data"].shape, g["rois"], g["argmax"], g["grad"]), g["dbottom"])
