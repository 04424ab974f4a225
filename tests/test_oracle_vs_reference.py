"""Pin the oracle on the reference itself (dev container only: needs /root/reference).
Runs the reference's sources through oracle/ref_shim.py and compares on fresh seeds and at the
full reference shapes.  Skipped where the reference tree is absent (GPU box)."""
import numpy as np
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    r = ref_shim.load()
    r.config.cfg_from_file(r.yml)
    return r


@pytest.mark.parametrize("kw", [
    dict(res=0.1, zres=0.3, side_range=(-30., 30.), fwd_range=(0., 60), height_range=(-2, 0.4)),
    dict(res=0.1, zres=0.1, side_range=(-40., 40.), fwd_range=(0., 70.), height_range=(-2.0, 1.5)),
])
def test_raster_full_shapes(oracle, ref, kw):
    pts = oracle.synth_points(60000, seed=3)
    a = ref.read_lidar.point_cloud_2_top(pts, **kw)
    b = oracle.point_cloud_2_top(pts, **kw)
    assert a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.parametrize("key", ["TEST", "TRAIN"])
def test_proposal_layer_full_shape(oracle, ref, key):
    prob, deltas = oracle.synth_rpn_outputs(75, 75, seed=5)
    im_info = np.array([[601, 601, 1]], dtype=np.float32)
    want = ref.proposal_layer_tf.proposal_layer_3d(prob, deltas, im_info, oracle.KITTI_CALIB, key, [8, ], [1.0, 1.0])
    c = ref.cfg[key]
    cfg = {key: dict(RPN_PRE_NMS_TOP_N=c.RPN_PRE_NMS_TOP_N, RPN_POST_NMS_TOP_N=c.RPN_POST_NMS_TOP_N,
                     RPN_NMS_THRESH=c.RPN_NMS_THRESH, RPN_MIN_SIZE=c.RPN_MIN_SIZE)}
    got = oracle.proposal_layer_3d(prob, deltas, im_info, oracle.KITTI_CALIB, key, cfg=cfg)
    for w, g in zip(want, got):
        assert np.array_equal(w, g)


def test_nms_and_iou_random(oracle, ref):
    rng = np.random.default_rng(9)
    n = 3000
    x1, y1 = rng.integers(0, 500, n), rng.integers(0, 500, n)
    d = np.stack((x1, y1, x1 + rng.integers(5, 80, n), y1 + rng.integers(5, 80, n), rng.permutation(n) / n), 1)
    d = d.astype(np.float32)
    assert list(ref.cpu_nms.cpu_nms(d, 0.7)) == oracle.nms(d, 0.7)
    assert list(ref.cython_nms.nms(d, 0.3)) == oracle.nms(d, 0.3)
    b = d[:500, :4].astype(np.float64)
    q = d[500:530, :4].astype(np.float64)
    assert np.array_equal(ref.cython_bbox.bbox_overlaps(b, q), oracle.bbox_overlaps(b, q))


@pytest.mark.parametrize("seed", [3, 11])
def test_target_layers_full_shape(oracle, ref, seed):
    """anchor_target_layer + proposal_target_layer_3d at the reference shape (75x75 map, 22 500 anchors), both sides
    drawing from numpy's global RandomState with the same seed."""
    gt_bv, gt_3d, gt_cnr = oracle.synth_gt(6, seed=seed)
    cls = np.zeros((1, 75, 75, 8), np.float32)
    im_info = np.array([[601, 601, 1]], dtype=np.float32)
    np.random.seed(seed)
    want = ref.anchor_target_layer_tf.anchor_target_layer(cls, gt_bv, gt_3d, im_info, [8, ], [1.0, 1.0])
    np.random.seed(seed)
    got = oracle.anchor_target_layer(cls, gt_bv, gt_3d, im_info)
    for w, g in zip(want, got):
        assert w.dtype == g.dtype and np.array_equal(w, g)
    prob, deltas = oracle.synth_rpn_outputs(75, 75, seed=seed + 40)
    rois_bv, _, rois_3d = ref.proposal_layer_tf.proposal_layer_3d(prob, deltas, im_info, oracle.KITTI_CALIB, "TRAIN",
                                                                 [8, ], [1.0, 1.0])
    np.random.seed(seed + 1)
    want = ref.proposal_target_layer_tf.proposal_target_layer_3d(rois_bv, rois_3d, gt_bv, gt_3d, gt_cnr,
                                                                 oracle.KITTI_CALIB, 2)
    np.random.seed(seed + 1)
    got = oracle.proposal_target_layer_3d(rois_bv, rois_3d, gt_bv, gt_3d, gt_cnr, oracle.KITTI_CALIB, 2)
    for w, g in zip(want, got):
        assert w.dtype == g.dtype and np.array_equal(w, g)


def test_box_detect_postprocessing_helpers(oracle, ref):
    """corners_to_bv / bbox_transform_inv_cnr restatements (the tail of box_detect, test_mv.py:241-264) vs the reference."""
    rng = np.random.default_rng(2)
    p3d = np.column_stack((rng.uniform(2, 58, 200), rng.uniform(-28, 28, 200), rng.uniform(-2, 0, 200),
                           rng.uniform(1, 5, 200), rng.uniform(1, 3, 200), rng.uniform(1, 2, 200))).astype(np.float32)
    cnr = ref.transform.lidar_3d_to_corners(p3d)
    assert np.array_equal(cnr, oracle.lidar_3d_to_corners(p3d))
    both = np.hstack((cnr, cnr))
    assert np.array_equal(ref.transform.corners_to_bv(both), oracle.corners_to_bv(both))
    deltas = rng.normal(0, 0.1, (200, 48)).astype(np.float32)
    assert np.array_equal(ref.bbox_transform.bbox_transform_inv_cnr(cnr, deltas), oracle.bbox_transform_inv_cnr(cnr, deltas))


def test_roi_pool_restatement_vs_reference_op(oracle):
    """The reference's own RoiPoolOp / RoiPoolGradOp CPU kernels (roi_pooling_op.cc compiled unmodified against the
    stand-in TF headers, oracle/build_ref_roi_pool.py) against the C restatement the GPU tests use -- forward values,
    arg-max and the backward gather, at toy and at the reference's 75x75x512 / 46x155x512 shapes."""
    from oracle import ref_roi_pool

    assert ref_roi_pool.available()
    rng = np.random.default_rng(0)
    for (B, H, W, C, R) in [(1, 9, 11, 4, 6), (2, 20, 31, 7, 40), (1, 75, 75, 512, 300), (1, 46, 155, 512, 64)]:
        data = rng.normal(size=(B, H, W, C)).astype(np.float32)
        data[0, 0] = 1.25
        x1, y1 = rng.integers(-60, W * 8, R), rng.integers(-60, H * 8, R)
        rois = np.stack((rng.integers(0, B, R), x1, y1, x1 + rng.integers(0, 300, R), y1 + rng.integers(0, 200, R)), 1).astype(np.float32)
        rois[0] = [0, -500, -500, -300, -300]
        rois[1] = [0, 40, 40, 8, 8]
        rois[2, 1:] += 0.5
        wt, wa = ref_roi_pool.roi_pool_forward(data, rois)
        ot, oa = oracle.roi_pool_fwd(data, rois)
        assert np.array_equal(wt, ot) and np.array_equal(wa, oa)
        if C <= 8:   # the reference's backward is O(H*W*C*R): small cases only
            g = rng.normal(size=wt.shape).astype(np.float32)
            assert np.array_equal(ref_roi_pool.roi_pool_backward(data, rois, wa, g), oracle.roi_pool_bwd(data.shape, rois, oa, g))
