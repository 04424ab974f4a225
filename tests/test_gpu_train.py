"""Training path on the GPU against the CPU oracle: target layers (bit-exact index sets), one full train step
(losses and every parameter gradient within 1e-3 of max|ref|, the north-star float tolerance), Adam update.

Gradient parity is stated on IDENTICAL DISCRETE DECISIONS: sampled anchors / rois, ReLU on-off patterns, 2x2-pool
winners and ROI-pool arg-max are taken from the GPU run and teacher-forced into the oracle.  Without that, a ReLU whose
pre-activation lies within the forward tolerance of zero (measured: 0-5 of ~2e5 units per layer) flips between the two
runs and moves the gradient below it by O(1/sqrt(#units)) ~ 3e-3 -- a property of the function, not of a kernel
(measured with tools/debug_train_grads.py; the fp32 and fp64 oracles agree with each other to 1e-6)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GEOM_KW = dict(x_min=0, x_max=16, y_min=-8, y_max=8, res=0.1)
RASTER_KW = dict(res=0.1, zres=0.3, side_range=(-8., 8.), fwd_range=(0., 16.), height_range=(-2, 0.4))


def _relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _ulp_diff(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


@pytest.fixture()
def train_cfg():
    from mv3d_tf_b200.fast_rcnn.config import cfg, cfg_from_end2end_yml

    cfg_from_end2end_yml()
    cfg.USE_GPU_NMS = False
    return cfg


def test_anchor_target_layer_golden_and_oracle(oracle, golden_dir, train_cfg):
    from mv3d_tf_b200.rpn_msr.anchor_target_layer_tf import anchor_target_layer

    g = np.load(os.path.join(golden_dir, "targets.npz"))
    hf, wf = int(g["hf"]), int(g["wf"])
    cls = np.zeros((1, hf, wf, 8), np.float32)
    np.random.seed(int(g["seed"]))
    labels, targets, anchors, anchors_3d = anchor_target_layer(cls, g["gt_bv"], g["gt_3d"], g["im_info"], [8, ], [1.0, 1.0])
    assert np.array_equal(labels, g["at_labels"])                      # index sets: bit-exact vs the reference
    assert _ulp_diff(targets, g["at_targets"]).max() <= 1              # float64 log then cast: <= 1 ulp
    assert (targets != g["at_targets"]).mean() < 1e-3
    assert np.array_equal(anchors, g["at_anchors"]) and np.array_equal(anchors_3d, g["at_anchors_3d"])
    # reference shape, fresh seeds, against the oracle
    for seed in (5, 8):
        gt_bv, gt_3d, _ = oracle.synth_gt(7, seed=seed)
        cls = np.zeros((1, 75, 75, 8), np.float32)
        info = np.array([[601, 601, 1]], np.float32)
        np.random.seed(seed)
        want = oracle.anchor_target_layer(cls, gt_bv, gt_3d, info)
        np.random.seed(seed)
        got = anchor_target_layer(cls, gt_bv, gt_3d, info, [8, ], [1.0, 1.0])
        assert np.array_equal(want[0], got[0])
        assert _ulp_diff(want[1], got[1]).max() <= 1
        assert np.array_equal(want[2], got[2]) and np.array_equal(want[3], got[3])


def test_proposal_target_layer_golden_and_oracle(oracle, golden_dir, train_cfg):
    from mv3d_tf_b200.rpn_msr.proposal_target_layer_tf import proposal_target_layer_3d

    g = np.load(os.path.join(golden_dir, "targets.npz"))
    np.random.seed(int(g["seed"]) + 1)
    out = proposal_target_layer_3d(g["rois_bv"], g["rois_3d"], g["gt_bv"], g["gt_3d"], g["gt_cnr"], g["calib"], 2)
    for got, key in zip(out, ("pt_rois_bv", "pt_rois_img", "pt_labels", "pt_targets", "pt_rois_3d")):
        assert got.dtype == g[key].dtype and got.shape == g[key].shape, key
        assert np.array_equal(got, g[key]), key                        # bit-exact vs the reference's own output
    for seed in (5, 8):
        gt_bv, gt_3d, gt_cnr = oracle.synth_gt(6, seed=seed)
        prob, deltas = oracle.synth_rpn_outputs(75, 75, seed=seed + 40)
        info = np.array([[601, 601, 1]], np.float32)
        cfgd = {"TRAIN": dict(RPN_PRE_NMS_TOP_N=12000, RPN_POST_NMS_TOP_N=2000, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=5)}
        with np.errstate(all="ignore"):
            rb, _, r3 = oracle.proposal_layer_3d(prob, deltas, info, oracle.KITTI_CALIB, "TRAIN", cfg=cfgd)
        np.random.seed(seed)
        want = oracle.proposal_target_layer_3d(rb, r3, gt_bv, gt_3d, gt_cnr, oracle.KITTI_CALIB, 2)
        np.random.seed(seed)
        got = proposal_target_layer_3d(rb, r3, gt_bv, gt_3d, gt_cnr, oracle.KITTI_CALIB, 2)
        for w, h in zip(want, got):
            assert np.array_equal(w, h)


def _small_calib(oracle):
    """KITTI calib with P2 scaled so that the synthetic 16 m x 16 m scene projects into the 64 x 256 test image."""
    c = np.array(oracle.KITTI_CALIB, dtype=np.float32).copy()
    p2 = c[0].reshape(3, 4)
    p2[0] *= 0.2
    p2[1] *= 0.17
    c[0] = p2.reshape(-1)
    return c


def _collect_gates(net, vals, b, roi_slice):
    """The discrete decisions the GPU run took for frame b: ReLU on/off per unit, 2x2 pool winners, ROI-pool arg-max."""
    from mv3d_tf_b200 import kernels as K
    from mv3d_tf_b200.fast_rcnn.train_mv import _node

    relu, pool, roi = {}, {}, {}
    for n in net._program:
        if n.kind == "conv" and n.name not in ("rpn_cls_score", "rpn_bbox_pred"):
            relu[n.name] = (K.unpad_nhwc(vals[n].pad)[b:b + 1] > 0).cpu().numpy()
        elif n.kind == "fc" and n.attrs.get("relu", True):
            v = vals[n]
            relu[n.name] = (v.hi[roi_slice, :n.channels].float() > 0).cpu().numpy()
        elif n.kind == "max_pool":
            x = K.unpad_nhwc(vals[n.inputs[0]].pad)[b:b + 1].permute(0, 3, 1, 2).cpu().double()
            _, idx = torch.nn.functional.max_pool2d(x, 2, 2, return_indices=True)
            pool[n.name] = idx.numpy()
        elif n.kind == "roi_pool":
            roi[n.name] = vals[n].extra["argmax"][roi_slice].cpu().numpy()
    return dict(relu=relu, pool=pool, roi=roi)


def _make_problem(oracle, B, seed=0):
    from mv3d_tf_b200.networks.factory import get_network
    from mv3d_tf_b200.utils.transform import BevGeometry

    geom = BevGeometry(**GEOM_KW)
    ogeom = oracle.BevGeometry(**GEOM_KW)
    net = get_network("MV3D_train", bv_channels=9, precise=True, geometry=geom, img_size=(64, 256))
    net.init_weights(seed=7, mode="he")
    frames = []
    for b in range(B):
        pts = oracle.synth_points(30000, seed=seed + b)
        pts[:, 0] *= 0.2
        pts[:, 1] *= 0.17
        bv = oracle.point_cloud_2_top(pts, **RASTER_KW)
        rng = np.random.default_rng(seed + 10 + b)
        img = rng.normal(0, 50, (64, 256, 3)).astype(np.float32)
        gt = oracle.synth_gt(4, seed=seed + 20 + b, geom=ogeom)
        frames.append(dict(bv=bv, img=img, gt=gt))
    blobs = dict(lidar_bv_data=np.stack([f["bv"] for f in frames]), image_data=np.stack([f["img"] for f in frames]),
                 im_info=np.array([[161, 161, 1]], np.float32), gt_boxes_bv=[f["gt"][0] for f in frames],
                 gt_boxes_3d=[f["gt"][1] for f in frames], gt_boxes_corners=[f["gt"][2] for f in frames],
                 calib=_small_calib(oracle))
    return net, frames, blobs, ogeom


@pytest.mark.parametrize("B", [1, 2])
def test_train_step_gradients_match_oracle(oracle, train_cfg, B):
    from mv3d_tf_b200.fast_rcnn.train_mv import SolverWrapper, _node
    from oracle import net_oracle

    net, frames, blobs, ogeom = _make_problem(oracle, B)
    sw = SolverWrapper(network=net, keep_prob=1.0, lr=1e-3)
    params0 = sw.export_params()
    np.random.seed(3)
    loss = sw.train_step(blobs, keep_prob=1.0, apply_update=False)
    torch.cuda.synchronize()
    vals = net.last_vals
    grads = sw.export_grads()
    rd = vals[_node(net, "roi_data_3d")].extra
    ad = vals[_node(net, "rpn_data")].extra
    counts = rd["frame_counts"].cpu().numpy()
    offs = np.concatenate(([0], np.cumsum(counts)))
    tot_losses = np.zeros(4)
    ref_grads = None
    for b in range(B):
        sl = slice(int(offs[b]), int(offs[b + 1]))
        rois_bv = rd["bv"][sl].cpu().numpy().copy()
        rois_img = rd["img"][sl].cpu().numpy().copy()
        rois_bv[:, 0] = 0
        rois_img[:, 0] = 0          # the oracle runs frame by frame (the reference is strictly batch 1)
        teacher = dict(rpn_data=(ad["labels"][b].cpu().numpy(), ad["targets"][b].cpu().numpy()),
                       roi_data=(rois_bv, rois_img, rd["labels"][sl].cpu().numpy(), rd["targets"][sl].cpu().numpy()))
        f = frames[b]
        losses, g, stage = net_oracle.train_forward_backward(f["bv"][None], f["img"][None], blobs["im_info"],
                                                              blobs["calib"], *f["gt"], params0, geom=ogeom,
                                                              teacher=teacher, gates=_collect_gates(net, vals, b, sl))
        tot_losses += np.array([losses["rpn_loss_cls"], losses["rpn_loss_box"], losses["loss_cls"], losses["loss_box"]]) / B
        if ref_grads is None:
            ref_grads = {k: {kk: vv / B for kk, vv in v.items()} for k, v in g.items()}
        else:
            for k in g:
                for kk in g[k]:
                    ref_grads[k][kk] += g[k][kk] / B
    got_losses = loss.cpu().numpy()
    assert np.allclose(got_losses, tot_losses, rtol=1e-3, atol=1e-5), (got_losses, tot_losses)
    worst = {}
    for k in ref_grads:
        for kk in ref_grads[k]:
            if np.abs(ref_grads[k][kk]).max() < 1e-12:
                assert np.abs(grads[k][kk]).max() < 1e-9
                continue
            worst[k + "/" + kk] = _relerr(grads[k][kk], ref_grads[k][kk])
    bad = {k: v for k, v in worst.items() if v > 1e-3}
    assert not bad, bad


def test_train_step_discrete_stages_and_update(oracle, train_cfg):
    """B = 1, no teacher forcing: the target layers see the GPU's own RPN outputs; their outputs must equal the oracle's
    on those same inputs with the same seed; then two Adam steps must reduce nothing to NaN and move every parameter."""
    from mv3d_tf_b200.fast_rcnn.train_mv import SolverWrapper, _node
    from oracle import net_oracle

    net, frames, blobs, ogeom = _make_problem(oracle, 1, seed=5)
    sw = SolverWrapper(network=net, keep_prob=1.0, lr=1e-4)
    theta0 = sw.theta.clone()
    np.random.seed(11)
    sw.train_step(blobs, keep_prob=1.0, apply_update=False)
    torch.cuda.synchronize()
    vals = net.last_vals
    score = vals[_node(net, "rpn_cls_score")].dense.cpu().numpy()
    prob = vals[_node(net, "rpn_cls_prob_reshape")].dense.cpu().numpy()
    bbox = vals[_node(net, "rpn_bbox_pred")].dense.cpu().numpy()
    gt_bv, gt_3d, gt_cnr = frames[0]["gt"]
    np.random.seed(11)
    labels, targets, _, _ = oracle.anchor_target_layer(score, gt_bv, gt_3d, blobs["im_info"], geom=ogeom)
    cfgd = {"TRAIN": dict(RPN_PRE_NMS_TOP_N=12000, RPN_POST_NMS_TOP_N=2000, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=5)}
    with np.errstate(all="ignore"):
        rb, _, r3 = oracle.proposal_layer_3d(prob, bbox, blobs["im_info"], blobs["calib"], "TRAIN", cfg=cfgd, geom=ogeom,
                                             img_size=(64, 256))
    want = oracle.proposal_target_layer_3d(rb, r3, gt_bv, gt_3d, gt_cnr, blobs["calib"], 2)
    ad = vals[_node(net, "rpn_data")].extra
    rd = vals[_node(net, "roi_data_3d")].extra
    assert np.array_equal(ad["labels"][0].cpu().numpy(), labels)
    assert _ulp_diff(ad["targets"][0].cpu().numpy(), targets).max() <= 1
    assert np.array_equal(rd["bv"].cpu().numpy(), want[0])
    assert np.array_equal(rd["img"].cpu().numpy(), want[1])
    assert np.array_equal(rd["labels"].cpu().numpy().reshape(-1, 1), want[2])
    # corner targets inherit the <= 2 ulp float32 `exp` difference of the proposal layer's l/w/h
    assert np.allclose(rd["targets"].cpu().numpy(), want[3], rtol=1e-4, atol=1e-6)
    assert np.array_equal(rd["targets"].cpu().numpy() != 0, want[3] != 0)
    # Adam: against the TF formula applied to the exported gradients
    np.random.seed(11)
    l1 = sw.train_step(blobs, keep_prob=1.0).clone()
    torch.cuda.synchronize()
    g = sw.grad.clone()   # the gradient this step applied (Adam does not modify it)
    lr_t = 1e-4 * np.sqrt(1 - 0.999) / (1 - 0.9)
    gd = g.double()
    exp = theta0.double() - lr_t * (0.1 * gd) / ((0.001 * gd * gd).sqrt() + 1e-8)
    assert float((sw.theta.double() - exp).abs().max()) < 1e-6
    l2 = sw.train_step(blobs, keep_prob=0.5).clone()
    torch.cuda.synchronize()
    assert bool(torch.isfinite(l1).all()) and bool(torch.isfinite(l2).all()) and bool(torch.isfinite(sw.theta).all())
