"""Front-view branch (this project's extension -- the reference has no FV, network.py:313-315) against its written
specification in oracle/mv3d_oracle.py: cylindrical raster bit-exact, FV rois bit-exact, three-view network within 1e-3."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = a.detach().cpu().double() if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a)).double()
    b = b.detach().cpu().double() if isinstance(b, torch.Tensor) else torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("n", [0, 1, 5000, 120000])
def test_fv_raster_bit_exact(oracle, n):
    from mv3d_tf_b200 import kernels as K
    from mv3d_tf_b200.utils.read_lidar import FvRasterizer

    pts = oracle.synth_points(max(n, 1), seed=3 + n)[:n]
    if n >= 5000:   # duplicates of one cell (last writer wins) and points behind / above / below the fan
        pts[100:200, :3] = pts[50, :3]
        pts[300, 0] = -5.0
        pts[301, 2] = 50.0
    r = FvRasterizer()
    d = torch.from_numpy(np.ascontiguousarray(pts)).cuda()
    want = oracle.point_cloud_2_front(pts)
    got = r(d).cpu().numpy()
    assert got.shape == (64, 512, 3) and np.array_equal(got, want)
    pad = r.to_pad(d, precise=True)
    w = torch.from_numpy(want)
    hi = w.bfloat16().float()
    lo = (w - hi).bfloat16().float()
    assert torch.equal(K.unpad_nhwc(pad)[0].cpu(), hi + lo)      # the trunk input is the bf16 hi/lo pair of the map
    assert float(pad.hi[:, :, 0].abs().max()) == 0 and float(pad.hi[:, 64].abs().max()) == 0


def test_rois_to_fv_bit_exact(oracle):
    from mv3d_tf_b200._lib import check, current_stream, lib, ptr
    from mv3d_tf_b200.utils.transform import FV_GEOMETRY

    rng = np.random.default_rng(5)
    R = 500
    p = np.column_stack((np.zeros(R), rng.uniform(-5, 70, R), rng.uniform(-40, 40, R), rng.uniform(-3, 1, R),
                         rng.uniform(0.5, 6, R), rng.uniform(0.5, 6, R), rng.uniform(1, 2.5, R))).astype(np.float32)
    p[7, 4] = np.nan
    p[8, 1] = np.inf
    d = torch.from_numpy(p).cuda()
    out = torch.empty((R, 5), dtype=torch.float32, device="cuda")
    num = torch.tensor([R - 10], dtype=torch.int32, device="cuda")
    H, W, t0, dt, p1, dp = FV_GEOMETRY.c_args()
    check(lib().mv3d_rois_to_fv(ptr(d), R, ptr(num), H, W, t0, dt, p1, dp, ptr(out), current_stream()), "mv3d_rois_to_fv")
    got = out.cpu().numpy()
    want = oracle.lidar_3d_to_fv(p[:, 1:7])
    assert np.array_equal(got[:R - 10, 1:], want[:R - 10])
    assert not got[R - 10:].any()


def test_three_view_network_vs_oracle(oracle):
    from mv3d_tf_b200.fast_rcnn.config import cfg, cfg_from_end2end_yml
    from mv3d_tf_b200.networks.factory import get_network
    from mv3d_tf_b200.utils.read_lidar import FvRasterizer
    from oracle import net_oracle

    cfg_from_end2end_yml()
    cfg.USE_GPU_NMS = False
    net = get_network("MV3D_test", bv_channels=9, precise=True, fv=True)
    net.init_weights(seed=7, mode="he")
    assert net.layers["roi_data_fv"] is not None
    pts = oracle.synth_points(40000, seed=9)
    pts[:, 0] *= 0.2
    pts[:, 1] *= 0.17
    bv = oracle.point_cloud_2_top(pts, 0.1, 0.3, (-8., 8.), (0., 16.), (-2, 0.4))[None]
    fv = oracle.point_cloud_2_front(pts)[None]
    rng = np.random.default_rng(12)
    img = rng.normal(0, 50, (1, 96, 320, 3)).astype(np.float32)
    im_info = np.array([[161, 161, 1]], np.float32)
    feed = {net.lidar_bv_data: bv, net.lidar_fv_data: FvRasterizer().to_pad(torch.from_numpy(pts).cuda()),
            net.image_data: img, net.im_info: im_info, net.calib: oracle.KITTI_CALIB}
    names = ["conv5_3", "conv5_3_2", "conv5_3_3", "pool_5_3", "cls_prob", "bbox_pred", "roi_data_bv", "roi_data_img",
             "roi_data_fv"]
    out = dict(zip(names, net.run([net.get_output(n) for n in names], feed)))
    rois3d = net.run([net.get_output("rois")], feed)[0]["p3d"]
    num = int(net.last_num_rois.item())
    assert num > 0
    params = {k: {kk: vv.cpu().numpy() for kk, vv in v.items()} for k, v in net.params.items()}
    ref_trunk = net_oracle.trunk(fv, params, "_3")
    assert _rel(out["conv5_3_3"], ref_trunk) < 1e-3
    t = net_oracle.mv3d_test_forward(bv, img, im_info, oracle.KITTI_CALIB, params, fv=fv, teacher=dict(
        conv5_3=out["conv5_3"].cpu().numpy(), conv5_3_2=out["conv5_3_2"].cpu().numpy(),
        conv5_3_3=out["conv5_3_3"].cpu().numpy(),
        rois=(out["roi_data_bv"][:num].cpu().numpy(), out["roi_data_img"][:num].cpu().numpy(), rois3d[:num].cpu().numpy())))
    assert np.array_equal(out["roi_data_fv"][:num].cpu().numpy(), t["rois_fv"])
    assert np.array_equal(out["pool_5_3"][:num].cpu().numpy(), t["pool_5_3"])
    assert _rel(out["bbox_pred"][:num], t["bbox_pred"]) < 1e-3
    assert float((out["cls_prob"][:num].cpu() - t["cls_prob"]).abs().max()) < 1e-4
    # the two-view network still returns None for the FV transform, as the reference does
    net2 = get_network("MV3D_test", bv_channels=9, precise=True)
    (net2.feed("rois").proposal_transform(target="fv", name="roi_data_fv_probe"))
    assert net2.layers["roi_data_fv_probe"] is None
