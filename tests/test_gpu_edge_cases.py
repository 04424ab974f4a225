"""Edge cases of SURVEY A11 / A5 and the round-2 kernels, on the GPU against the CPU oracle:
tied scores (pinned rule: two stable `argsort()[::-1]`, proposal_layer_tf.py:161 + cpu_nms.pyx:25), zero-overlap
ground truth (anchor_target_layer_tf.py:123), exp overflow in bbox_transform_inv_3d, empty survivor sets
(nms_wrapper.py:16-17), clustered boxes through the super-block keep chain, and the fused multi-view ROI pool with
in-kernel projection (BEV / image / FV rectangles from the 3-D proposals) against projection + pooling by the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cfg():
    from mv3d_tf_b200.fast_rcnn.config import cfg, cfg_from_end2end_yml

    cfg_from_end2end_yml()
    cfg.USE_GPU_NMS = False
    return cfg


def _layer_vs_oracle(oracle, prob, deltas, key="TEST", **over):
    from mv3d_tf_b200.rpn_msr.proposal_layer_tf import ProposalLayer3D

    cfg = _cfg()
    hf, wf = prob.shape[1:3]
    c = cfg[key]
    kw = dict(pre_nms_top_n=over.get("pre", c.RPN_PRE_NMS_TOP_N), post_nms_top_n=over.get("post", c.RPN_POST_NMS_TOP_N))
    layer = ProposalLayer3D(hf, wf, key, 8, (601, 601, 1), **kw)
    out = layer(torch.from_numpy(prob[0]).cuda(), torch.from_numpy(deltas[0]).cuda(), oracle.KITTI_CALIB)
    n = int(out["num"].item())
    ocfg = {key: dict(RPN_PRE_NMS_TOP_N=kw["pre_nms_top_n"], RPN_POST_NMS_TOP_N=kw["post_nms_top_n"],
                      RPN_NMS_THRESH=c.RPN_NMS_THRESH, RPN_MIN_SIZE=c.RPN_MIN_SIZE)}
    with np.errstate(all="ignore"):
        (bv, img, p3d), st = oracle.proposal_layer_3d(prob, deltas, np.array([[601, 601, 1]], np.float32),
                                                      oracle.KITTI_CALIB, key, cfg=ocfg, return_stages=True)
    assert n == bv.shape[0], (n, bv.shape)
    assert np.array_equal(out["anchor"][:n].cpu().numpy(), st["anchor_index"].astype(np.int32))
    assert np.array_equal(out["bv"][:n].cpu().numpy(), bv)
    assert np.array_equal(out["img"][:n].cpu().numpy(), img)
    assert float(out["bv"][n:].abs().sum()) == 0
    return n


@pytest.mark.parametrize("levels,pre", [(7, 6000), (64, 6000), (3, 500), (1, 6000), (1000, 2000)])
def test_proposal_layer_tied_scores(oracle, levels, pre):
    """Scores quantised to a few distinct values: large tie groups, some straddling the pre-NMS top-N cut.  The device
    ranking must reproduce the oracle's order (ties higher-index-first at the cut, mirrored inside NMS)."""
    prob, deltas = oracle.synth_rpn_outputs(40, 37, seed=11)
    rng = np.random.default_rng(levels)
    fg = (rng.integers(0, levels, prob[..., 1::2].shape).astype(np.float32) + 1) / np.float32(levels + 1)
    prob = prob.copy()
    prob[..., 1::2] = fg
    prob[..., 0::2] = 1 - fg
    _layer_vs_oracle(oracle, prob, deltas, "TEST", pre=pre, post=300)
    _layer_vs_oracle(oracle, prob, deltas, "TRAIN", pre=pre, post=2000)


def test_proposal_layer_exp_overflow_and_empty(oracle):
    """A11 (iii): dl/dw/dh large enough for exp -> inf (then NaN through `//`): the same anchors must be dropped.
    A11 (iv): every proposal filtered out -> zero rows, count 0."""
    prob, deltas = oracle.synth_rpn_outputs(30, 33, seed=5)
    d = deltas.copy().reshape(-1, 6)
    rng = np.random.default_rng(2)
    rows = rng.choice(d.shape[0], 400, replace=False)
    d[rows[:150], 3] = 95.0          # exp overflows float32 and the double -> float cast: inf
    d[rows[150:300], 4] = 800.0      # overflows in double as well
    d[rows[300:], 5] = -800.0        # underflow: zero height
    n = _layer_vs_oracle(oracle, prob, d.reshape(deltas.shape), "TEST")
    assert n > 0
    d2 = np.full_like(d, 0.0)
    d2[:, 3:5] = -30.0               # every box shrinks below RPN_MIN_SIZE -> nothing survives the size filter
    n = _layer_vs_oracle(oracle, prob, d2.reshape(deltas.shape), "TEST")
    assert n == 0


def test_anchor_targets_zero_overlap_gt(oracle):
    """A11 (i): a GT box that overlaps no inside anchor has gt_max_overlaps == 0, so `overlaps == gt_max_overlaps`
    marks EVERY zero-overlap anchor as foreground before subsampling (py-faster-rcnn behaviour, kept)."""
    from mv3d_tf_b200.rpn_msr.anchor_target_layer_tf import anchor_target_layer

    _cfg()
    gt_bv, gt_3d, _ = oracle.synth_gt(3, seed=9)
    gt_bv = gt_bv.copy()
    gt_bv[1, :4] = [5000, 5000, 5040, 5016]      # far outside the 601x601 grid: IoU 0 with every anchor
    cls = np.zeros((1, 75, 75, 8), np.float32)
    info = np.array([[601, 601, 1]], np.float32)
    for seed in (1, 2):
        np.random.seed(seed)
        want = oracle.anchor_target_layer(cls, gt_bv, gt_3d, info)
        np.random.seed(seed)
        got = anchor_target_layer(cls, gt_bv, gt_3d, info, [8, ], [1.0, 1.0])
        assert np.array_equal(want[0], got[0])
        assert np.array_equal(want[2], got[2]) and np.array_equal(want[3], got[3])
        a = np.ascontiguousarray(want[1], np.float32).view(np.int32).astype(np.int64)
        b = np.ascontiguousarray(got[1], np.float32).view(np.int32).astype(np.int64)
        assert np.abs(a - b).max() <= 1


@pytest.mark.parametrize("n,clusters,thresh", [(6000, 40, 0.7), (12000, 300, 0.7), (3000, 3, 0.5), (1500, 1500, 0.7)])
def test_nms_clustered_boxes(oracle, n, clusters, thresh):
    """Heavily overlapping boxes (few survivors per 1024-candidate super-block) and early stop, both rules."""
    from mv3d_tf_b200.nms.gpu_nms import cpu_nms, gpu_nms, nms_device

    rng = np.random.default_rng(n + clusters)
    cx, cy = rng.integers(40, 560, clusters), rng.integers(40, 560, clusters)
    k = rng.integers(0, clusters, n)
    x1 = cx[k] + rng.integers(-6, 7, n)
    y1 = cy[k] + rng.integers(-6, 7, n)
    d = np.stack((x1, y1, x1 + rng.integers(30, 44, n), y1 + rng.integers(12, 20, n), rng.permutation(n) / n), 1).astype(np.float32)
    want = oracle.nms(d, thresh, "ge")
    assert cpu_nms(d, thresh) == want
    assert gpu_nms(d, thresh) == oracle.nms(d, thresh, "gt")
    order = np.argsort(d[:, 4], kind="stable")[::-1]
    boxes = torch.from_numpy(np.ascontiguousarray(d[order, :4])).cuda()
    for cap in (1, 17, 300, 2000, len(want)):       # <= 2048: the lazy cluster kernel; above: mask + keep chain
        keep, num = nms_device(boxes, thresh, True, max_keep=cap)
        m = int(num.item())
        assert m == min(cap, len(want))
        assert list(order[keep[:m].cpu().numpy()]) == want[:m]
    want_gt = oracle.nms(d, thresh, "gt")
    keep, num = nms_device(boxes, thresh, False, max_keep=300)
    m = int(num.item())
    assert m == min(300, len(want_gt)) and list(order[keep[:m].cpu().numpy()]) == want_gt[:m]


@pytest.mark.parametrize("n", [1, 2, 511, 512, 513, 1025, 6000, 12000])
@pytest.mark.parametrize("spread", [560, 120])
def test_nms_lazy_kernel_equals_full_chain(oracle, n, spread):
    """mv3d_nms with 0 < max_keep <= 2048 (nms_lazy_kernel: candidates tested against kept boxes only, 8-CTA cluster)
    returns the prefix of the full greedy chain, for sparse (nothing suppressed: the block fast path) and dense inputs,
    block-boundary sizes, and a device-side box count."""
    from mv3d_tf_b200._lib import check, current_stream, lib, ptr
    from mv3d_tf_b200.nms.gpu_nms import nms_device

    rng = np.random.default_rng(7 * n + spread)
    x1, y1 = rng.integers(0, spread, n), rng.integers(0, spread, n)
    b = np.stack((x1, y1, x1 + rng.integers(4, 70, n), y1 + rng.integers(4, 70, n)), 1).astype(np.float32)
    sc = (np.arange(n)[::-1] / max(n, 1)).astype(np.float32)          # already sorted
    full = oracle.nms(np.hstack((b, sc[:, None])), 0.7, "ge")
    boxes = torch.from_numpy(b).cuda()
    for cap in (1, 300, 2000, 2048):
        keep, num = nms_device(boxes, 0.7, True, max_keep=cap)
        m = int(num.item())
        assert m == min(cap, len(full)) and keep[:m].cpu().tolist() == full[:m], (n, spread, cap)
    # device-side count smaller than the capacity (the proposal layer's form)
    n_dev = max(1, n - 37)
    full_dev = oracle.nms(np.hstack((b[:n_dev], sc[:n_dev, None])), 0.7, "ge")
    L = lib()
    keep = torch.empty(n, dtype=torch.int32, device="cuda")
    num = torch.zeros(1, dtype=torch.int32, device="cuda")
    d_n = torch.tensor([n_dev], dtype=torch.int32, device="cuda")
    ws = torch.empty(L.mv3d_nms_workspace_bytes(n), dtype=torch.uint8, device="cuda")
    check(L.mv3d_nms(ptr(boxes), n, 4, ptr(d_n), 0.7, 1, 300, ptr(keep), ptr(num), ptr(ws), ws.numel(), current_stream()), "mv3d_nms")
    m = int(num.item())
    assert m == min(300, len(full_dev)) and keep[:m].cpu().tolist() == full_dev[:m]


def _fused_pool(views, rois_3d, proj, R, num, Cc):
    from mv3d_tf_b200._lib import check, current_stream, lib, ptr

    check(lib().mv3d_roi_pool_fused(views, len(views), ptr(rois_3d), C.byref(proj), R, ptr(num), Cc, 7, 7,
                                    current_stream()), "mv3d_roi_pool_fused")


def test_roi_pool_fused_projects_and_pools_three_views(oracle):
    """mv3d_roi_pool_fused: rectangles projected in the kernel == the oracle's lidar_3d_to_bv + clip / lidar_cnr_to_img /
    lidar_3d_to_fv; pooled values + arg-max == the oracle's RoiPool on those rectangles; rows past the count are zero.
    Includes boxes behind the camera (inf/NaN -> INT_MIN image boxes), windows larger than one shared-memory stage
    (channel slices) and whole-map windows (direct reads)."""
    from mv3d_tf_b200._lib import ROI_BEV, ROI_FV, ROI_GIVEN, ROI_IMG, RoiProjection, RoiView, ptr
    from mv3d_tf_b200.utils.transform import CFG_GEOMETRY, FV_GEOMETRY, projection_matrix

    rng = np.random.default_rng(8)
    R, valid, Cc = 160, 150, 128
    p = np.column_stack((np.zeros(R), rng.uniform(2, 68, R), rng.uniform(-38, 38, R), rng.uniform(-2.5, 0.5, R),
                         rng.uniform(1, 6, R), rng.uniform(0.6, 3, R), rng.uniform(1, 2.2, R))).astype(np.float32)
    p[3, 1:4] = [-4.0, 1.0, -1.0]        # behind the camera
    p[4, 1] = 0.3                        # straddles the camera plane: corners with non-positive depth
    p[5, 4:6] = [60.0, 70.0]             # enormous box: whole-map windows
    p[6, 4:6] = [9.0, 7.0]               # ~120-cell BEV window: two channel slices
    p[7, 4:6] = [25.0, 20.0]             # ~800 cells: direct reads
    geom = oracle.CFG_GEOMETRY
    with np.errstate(all="ignore"):
        bv = oracle.clip_boxes(oracle.lidar_3d_to_bv(p[:, 1:7], geom), np.array([701, 801], np.float32))
        img = oracle.lidar_cnr_to_img(oracle.lidar_3d_to_corners(p[:, 1:7]), oracle.KITTI_CALIB[3], oracle.KITTI_CALIB[2],
                                      oracle.KITTI_CALIB[0]).astype(np.float32)
        fv = oracle.lidar_3d_to_fv(p[:, 1:7])
    maps = [rng.normal(size=(1, 87, 100, Cc)).astype(np.float32), rng.normal(size=(1, 46, 155, Cc)).astype(np.float32),
            rng.normal(size=(1, 8, 64, Cc)).astype(np.float32)]
    dev = [torch.from_numpy(m).cuda() for m in maps]
    tops = [torch.empty((R, 7, 7, Cc), device="cuda") for _ in range(3)]
    args = [torch.empty((R, 7, 7, Cc), dtype=torch.int32, device="cuda") for _ in range(3)]
    his = [torch.empty((R, 49 * Cc), dtype=torch.bfloat16, device="cuda") for _ in range(3)]
    los = [torch.empty_like(h) for h in his]
    outs = [torch.full((R, 5), -7.0, device="cuda") for _ in range(3)]
    views = (RoiView * 3)()
    for k, (m, src) in enumerate(zip(dev, (ROI_BEV, ROI_IMG, ROI_FV))):
        v = views[k]
        v.d_data, v.d_rois, v.height, v.width, v.spatial_scale = ptr(m), None, m.shape[1], m.shape[2], 0.125
        v.d_top, v.d_argmax, v.d_top_hi, v.d_top_lo = ptr(tops[k]), ptr(args[k]), ptr(his[k]), ptr(los[k])
        v.source, v.d_rois_out = src, ptr(outs[k])
    proj = RoiProjection()
    g = CFG_GEOMETRY
    proj.xn, proj.yn, proj.x_min, proj.y_min, proj.res = g.xn, g.yn, g.x_min, g.y_min, g.res
    proj.im_h, proj.im_w = 701.0, 801.0
    proj.h_proj = (C.c_float * 12)(*[float(x) for x in projection_matrix(oracle.KITTI_CALIB).reshape(-1)])
    proj.fv_h, proj.fv_w, proj.fv_theta_min, proj.fv_dtheta, proj.fv_phi_max, proj.fv_dphi = FV_GEOMETRY.c_args()
    num = torch.tensor([valid], dtype=torch.int32, device="cuda")
    _fused_pool(views, torch.from_numpy(p).cuda(), proj, R, num, Cc)
    torch.cuda.synchronize()
    for k, rect in enumerate((bv, img, fv)):
        want_rois = np.hstack((p[:, :1], rect)).astype(np.float32)
        got_rois = outs[k].cpu().numpy()
        assert np.array_equal(got_rois[:valid], want_rois[:valid]), k
        assert not got_rois[valid:].any()
        wt, wa = oracle.roi_pool_fwd(maps[k], want_rois[:valid])
        assert np.array_equal(tops[k][:valid].cpu().numpy(), wt), k
        assert np.array_equal(args[k][:valid].cpu().numpy(), wa), k
        assert not tops[k][valid:].any() and bool((args[k][valid:] == -1).all())
        hi, lo = his[k][:valid].float().cpu().numpy(), los[k][:valid].float().cpu().numpy()
        assert np.abs(hi + lo - wt.reshape(valid, -1)).max() <= 2.0 ** -15 * np.abs(wt).max()
    # the same kernel with GIVEN rectangles (the reference op's form) gives the same bits
    views2 = (RoiView * 1)()
    r_dev = torch.from_numpy(np.hstack((p[:, :1], img)).astype(np.float32)).cuda()
    top2 = torch.empty_like(tops[1])
    v = views2[0]
    v.d_data, v.d_rois, v.height, v.width, v.spatial_scale = ptr(dev[1]), ptr(r_dev), 46, 155, 0.125
    v.d_top, v.source = ptr(top2), ROI_GIVEN
    _fused_pool(views2, torch.from_numpy(p).cuda(), proj, R, num, Cc)
    assert torch.equal(top2, tops[1])


@pytest.mark.parametrize("fmt", [0, 1], ids=["bf16x2", "f16e5"])
def test_roi_pool_reads_pad_operand_planes(oracle, fmt):
    """The fused ROI pool fed from the PAD operand planes a conv writes (no dense float32 copy): values and arg-max equal
    the oracle's RoiPool on the operand RENDERING of the map (what Network.run returns for that layer), bit for bit."""
    from mv3d_tf_b200 import kernels as k
    from mv3d_tf_b200._lib import ROI_GIVEN, RoiView, check, current_stream, lib, ptr

    rng = np.random.default_rng(3 + fmt)
    H, W, Cc, R = 46, 155, 128, 90
    x = torch.from_numpy(rng.normal(size=(1, H, W, Cc)).astype(np.float32)).cuda()
    pad = k.pad_nhwc(x, precise=True, fmt=fmt)
    rendered = k.unpad_nhwc(pad).cpu().numpy()
    x1, y1 = rng.integers(-60, W * 8, R), rng.integers(-60, H * 8, R)
    rois = np.stack((np.zeros(R), x1, y1, x1 + rng.integers(0, 300, R), y1 + rng.integers(0, 200, R)), 1).astype(np.float32)
    rois[0] = [0, 0, 0, W * 8, H * 8]          # whole map: direct reads
    r_dev = torch.from_numpy(rois).cuda()
    top = torch.empty((R, 7, 7, Cc), device="cuda")
    arg = torch.empty((R, 7, 7, Cc), dtype=torch.int32, device="cuda")
    views = (RoiView * 1)()
    v = views[0]
    v.d_data, v.d_rois, v.height, v.width, v.spatial_scale = None, ptr(r_dev), H, W, 0.125
    v.d_top, v.d_argmax, v.source = ptr(top), ptr(arg), ROI_GIVEN
    v.d_pad_hi, v.d_pad_lo, v.pad_fmt, v.pad_c = ptr(pad.hi), ptr(pad.lo), fmt, pad.c_pad
    check(lib().mv3d_roi_pool_multiview(views, 1, R, None, Cc, 7, 7, current_stream()), "mv3d_roi_pool_multiview")
    wt, wa = oracle.roi_pool_fwd(rendered, rois)
    assert np.array_equal(top.cpu().numpy(), wt)
    assert np.array_equal(arg.cpu().numpy(), wa)


def test_roi_pool_emits_f16e5_fc_operand(oracle):
    """mv3d_roi_view.top_fmt = MV3D_FMT_F16E5: the pooled rows leave as fc6's 2-pass operand -- fp16 plane = fp16(top) bit
    for bit, byte plane = per 64-element chunk [e5m2(h) | e5m2((top - h) * 4096)] -- next to the exact float32 top."""
    from mv3d_tf_b200._lib import ROI_GIVEN, RoiView, check, current_stream, lib, ptr

    rng = np.random.default_rng(17)
    H, W, Cc, R = 40, 60, 128, 70
    x = torch.from_numpy(np.maximum(rng.normal(size=(1, H, W, Cc)), 0).astype(np.float32)).cuda()
    x1, y1 = rng.integers(-40, W * 8, R), rng.integers(-40, H * 8, R)
    rois = np.stack((np.zeros(R), x1, y1, x1 + rng.integers(0, 250, R), y1 + rng.integers(0, 200, R)), 1).astype(np.float32)
    r_dev = torch.from_numpy(rois).cuda()
    n_valid = torch.tensor([R - 5], dtype=torch.int32, device="cuda")     # rows past the count are zero-filled
    top = torch.empty((R, 7, 7, Cc), device="cuda")
    Kd = 49 * Cc
    hi = torch.full((R, Kd), -1, dtype=torch.int16, device="cuda")
    lo = torch.full((R, Kd), -1, dtype=torch.int16, device="cuda")
    views = (RoiView * 1)()
    v = views[0]
    v.d_data, v.d_rois, v.height, v.width, v.spatial_scale = ptr(x), ptr(r_dev), H, W, 0.125
    v.d_top, v.source, v.d_top_hi, v.d_top_lo, v.top_fmt = ptr(top), ROI_GIVEN, ptr(hi), ptr(lo), 1
    check(lib().mv3d_roi_pool_multiview(views, 1, R, ptr(n_valid), Cc, 7, 7, current_stream()), "mv3d_roi_pool_multiview")
    torch.cuda.synchronize()
    wt, _ = oracle.roi_pool_fwd(x.cpu().numpy(), rois[: R - 5])
    t = top.view(R, Kd)
    assert np.array_equal(t[: R - 5].cpu().numpy(), wt.reshape(R - 5, Kd))
    t = t.clone()
    t[R - 5:] = 0
    h = t.half()
    assert torch.equal(hi.view(torch.float16), h)
    planes = lo.view(torch.uint8).view(R, Kd // 64, 2, 64)
    h8 = planes[:, :, 0].reshape(R, Kd).view(torch.float8_e5m2).float()
    l8 = planes[:, :, 1].reshape(R, Kd).view(torch.float8_e5m2).float()
    assert torch.equal(h8, h.float().to(torch.float8_e5m2).float())
    resid = (t - h.float()) * 4096.0
    assert torch.equal(l8, resid.to(torch.float8_e5m2).float())
    assert float(((h.float() + l8 / 4096.0) - t).abs().max() / t.abs().max()) < 2.0 ** -13
