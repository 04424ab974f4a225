"""Float parity at the BASELINE shapes: BASELINE.json configs[1] -- 120 k-point LiDAR -> BEV
701x801x36, RGB 375x1242 (+ FV 64x512x3), the production tile predication (Wp = 802 / 1243, M = 563 004) -- and one
configs[2] train step (2 frames/GPU) against the CPU oracle ON THE SAME WEIGHTS.

Tolerance: 1e-3 relative (north_star) with the metric max|a-b| / max|b| per tensor; pooled maps are a
pure max/copy and must be bit-identical on identical inputs; proposals are exact on identical RPN outputs.
Reference: lib/networks/MV3D_test.py:33-123, MV3D_train.py:42-182, lib/fast_rcnn/train_mv.py:94-146."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BEV = dict(res=0.1, zres=0.1, side_range=(-40., 40.), fwd_range=(0., 70.), height_range=(-2.0, 1.5))   # 701x801x36
IMG_HW = (375, 1242)
PIXEL_MEANS = np.array([95.8814, 98.7743, 93.8549], np.float32)


def _frame(oracle, seed):
    pts = oracle.synth_points(120000, seed=1234 + seed)
    rng = np.random.default_rng(99 + seed)
    img = rng.integers(0, 256, (1, IMG_HW[0], IMG_HW[1], 3)).astype(np.float32) - PIXEL_MEANS
    return pts, img.astype(np.float32)


@pytest.mark.parametrize("views,mode", [(2, "mixed"), (2, "precise"), (3, "mixed"), (3, "precise")])
def test_configs1_full_shape_vs_oracle(oracle, views, mode):
    from oracle import parity
    from mv3d_tf_b200.fast_rcnn.config import cfg, cfg_from_end2end_yml
    from mv3d_tf_b200.networks.factory import get_network
    from mv3d_tf_b200.utils.read_lidar import BevRasterizer, FvRasterizer
    from mv3d_tf_b200.utils.transform import CFG_GEOMETRY

    cfg_from_end2end_yml()
    cfg.USE_GPU_NMS = False
    net = get_network("MV3D_test", bv_channels=36, precise=True, mixed=(mode == "mixed"), geometry=CFG_GEOMETRY,
                      fv=(views == 3))
    net.init_weights(seed=7, mode="he")
    pts, img = _frame(oracle, 0)
    im_info = np.array([[701, 801, 1]], np.float32)
    raster = BevRasterizer(**BEV)
    fvr = FvRasterizer(net.fv_geometry) if views == 3 else None
    got = parity.gpu_frame_outputs(net, raster, torch.from_numpy(pts).cuda(), img, im_info, oracle.KITTI_CALIB, fvr)
    assert got["num"] > 0
    bv = oracle.point_cloud_2_top(pts, **BEV)[None]
    assert bv.shape == (1, 701, 801, 36)
    fv = oracle.point_cloud_2_front(pts)[None] if views == 3 else None
    params = {k: {kk: vv.cpu().numpy() for kk, vv in v.items()} for k, v in net.params.items()}
    ocfg = {"TEST": dict(RPN_PRE_NMS_TOP_N=6000, RPN_POST_NMS_TOP_N=300, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=5)}
    errs, exact, prop = parity.oracle_frame_errors(got, params, bv, img, im_info, oracle.KITTI_CALIB,
                                                   oracle.CFG_GEOMETRY, cfg=ocfg, fv=fv)
    print("configs[1] %d views %s: errors %s exact %s proposals %s" % (views, mode, errs, exact, prop))
    for k, v in errs.items():
        assert v < parity.FLOAT_TOL, (k, v, errs)
    assert all(exact.values()), exact
    assert prop["iou90_match"] >= 0.9, prop


def test_configs2_full_shape_train_step_vs_oracle(oracle):
    """One configs[2] step: 2 frames of 701x801x36 + 375x1242, 6 GT cars each, RPN 12000/2000, 128 rois per frame;
    losses and every parameter gradient vs the torch-CPU autograd oracle on identical discrete decisions
    (tests/test_gpu_train.py states why the gates are teacher-forced)."""
    from test_gpu_train import _collect_gates, _relerr
    from mv3d_tf_b200.fast_rcnn.config import cfg, cfg_from_end2end_yml
    from mv3d_tf_b200.fast_rcnn.train_mv import SolverWrapper, _node
    from mv3d_tf_b200.networks.factory import get_network
    from mv3d_tf_b200.utils.transform import CFG_GEOMETRY
    from oracle import net_oracle

    cfg_from_end2end_yml()
    cfg.USE_GPU_NMS = False
    B = 2
    net = get_network("MV3D_train", bv_channels=36, precise=True, geometry=CFG_GEOMETRY)
    net.init_weights(seed=7, mode="he")
    frames = []
    for b in range(B):
        pts, img = _frame(oracle, 100 + b)
        frames.append(dict(bv=oracle.point_cloud_2_top(pts, **BEV), img=img[0],
                           gt=oracle.synth_gt(6, seed=500 + b, geom=oracle.CFG_GEOMETRY)))
    blobs = dict(lidar_bv_data=np.stack([f["bv"] for f in frames]), image_data=np.stack([f["img"] for f in frames]),
                 im_info=np.array([[701, 801, 1]], np.float32), gt_boxes_bv=[f["gt"][0] for f in frames],
                 gt_boxes_3d=[f["gt"][1] for f in frames], gt_boxes_corners=[f["gt"][2] for f in frames],
                 calib=oracle.KITTI_CALIB)
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
    tot = np.zeros(4)
    ref_grads = None
    for b in range(B):
        sl = slice(int(offs[b]), int(offs[b + 1]))
        rois_bv = rd["bv"][sl].cpu().numpy().copy()
        rois_img = rd["img"][sl].cpu().numpy().copy()
        rois_bv[:, 0] = 0
        rois_img[:, 0] = 0
        teacher = dict(rpn_data=(ad["labels"][b].cpu().numpy(), ad["targets"][b].cpu().numpy()),
                       roi_data=(rois_bv, rois_img, rd["labels"][sl].cpu().numpy(), rd["targets"][sl].cpu().numpy()))
        f = frames[b]
        losses, g, _ = net_oracle.train_forward_backward(f["bv"][None], f["img"][None], blobs["im_info"], blobs["calib"],
                                                         *f["gt"], params0, geom=oracle.CFG_GEOMETRY, teacher=teacher,
                                                         gates=_collect_gates(net, vals, b, sl))
        tot += np.array([losses["rpn_loss_cls"], losses["rpn_loss_box"], losses["loss_cls"], losses["loss_box"]]) / B
        if ref_grads is None:
            ref_grads = {k: {kk: vv / B for kk, vv in v.items()} for k, v in g.items()}
        else:
            for k in g:
                for kk in g[k]:
                    ref_grads[k][kk] += g[k][kk] / B
    got_losses = loss.cpu().numpy()
    print("configs[2] losses gpu %s oracle %s" % (got_losses, tot))
    assert np.allclose(got_losses, tot, rtol=1e-3, atol=1e-5), (got_losses, tot)
    worst = {}
    for k in ref_grads:
        for kk in ref_grads[k]:
            if np.abs(ref_grads[k][kk]).max() < 1e-12:
                assert np.abs(grads[k][kk]).max() < 1e-9
                continue
            worst[k + "/" + kk] = _relerr(grads[k][kk], ref_grads[k][kk])
    print("configs[2] worst gradient errors:", sorted(worst.items(), key=lambda kv: -kv[1])[:6])
    bad = {k: v for k, v in worst.items() if v > 1e-3}
    assert not bad, bad
