"""TEST INFRASTRUCTURE ONLY.  Shared by tests/test_gpu_fullconfig.py and bench.py's parity leg: one frame of MV3D_test on the GPU compared with the
CPU oracle (oracle/net_oracle.py, oracle/mv3d_oracle.py) on the SAME weights and inputs, at any shape.

Protocol = SURVEY Appendix C: (i) free-running float features per tensor, metric max|a-b| / max|b|, tolerance 1e-3 (the
north-star contract); (ii) teacher-forced tail: the oracle's ROI pool on the GPU's own conv5 maps and rois must be
bit-identical, the head within 1e-3; (iii) free-running proposals matched by IoU (reported).
This module is test infrastructure: it imports the oracle and must never be imported from mv3d_tf_b200/."""
from __future__ import annotations

import numpy as np
import torch

FLOAT_TOL = 1e-3   # north_star: "within 1e-3 relative for conv/pooled float features"


def rel(a, b):
    a = a.detach().cpu().double() if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a)).double()
    b = b.detach().cpu().double() if isinstance(b, torch.Tensor) else torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def iou_match_fraction(ref_boxes, got_boxes, thr=0.9):
    """Fraction of `ref_boxes` (n,4) that have a box in `got_boxes` with IoU >= thr (+1 pixel convention)."""
    if len(ref_boxes) == 0:
        return 1.0
    if len(got_boxes) == 0:
        return 0.0
    a = np.asarray(ref_boxes, np.float64)[:, None, :]
    b = np.asarray(got_boxes, np.float64)[None, :, :]
    iw = np.minimum(a[..., 2], b[..., 2]) - np.maximum(a[..., 0], b[..., 0]) + 1
    ih = np.minimum(a[..., 3], b[..., 3]) - np.maximum(a[..., 1], b[..., 1]) + 1
    inter = np.clip(iw, 0, None) * np.clip(ih, 0, None)
    area = lambda x: (x[..., 2] - x[..., 0] + 1) * (x[..., 3] - x[..., 1] + 1)
    iou = inter / (area(a) + area(b) - inter)
    return float((iou.max(axis=1) >= thr).mean())


# conv2_1 (not conv1_2): fetching a layer's dense output switches its pool fusion off, and the fused conv1_2 + pool1
# kernel is what production runs -- conv2_1 sits right behind it
FEATURES = ("conv2_1", "conv3_3", "conv5_3", "conv2_1_2", "conv3_3_2", "conv5_3_2", "rpn_bbox_pred")


def gpu_frame_outputs(net, raster, pts_dev, img, im_info, calib, fv_raster=None, features=FEATURES):
    """One eager frame on the GPU, the way FrameRunner feeds it (operand-format raster, dense FV map).  Returns a dict
    of CPU numpy arrays: float features, RPN outputs, the three roi blobs (sliced to the count), pooled maps, head."""
    from mv3d_tf_b200 import kernels as K

    fmt = K.FMT_F16E5 if (getattr(net, "mixed", False) and raster.g["C"] > 16) else K.FMT_BF16X2
    bv = raster.to_pad(pts_dev, precise=net.precise, fmt=fmt)
    feed = {net.lidar_bv_data: bv, net.image_data: img, net.im_info: im_info, net.calib: calib}
    views = 3 if getattr(net, "with_fv", False) else 2
    names = list(features) + ["rpn_cls_prob_reshape", "cls_prob", "bbox_pred", "roi_data_bv", "roi_data_img", "pool_5",
                              "pool_5_2"]
    if views == 3:
        feed[net.lidar_fv_data] = fv_raster(pts_dev)[None]
        names += ["conv5_3_3", "roi_data_fv", "pool_5_3"]
    vals = net.run([net.get_output(n) for n in names] + [net.get_output("rois")], feed)
    torch.cuda.synchronize()
    num = int(net.last_num_rois.item())
    out = {}
    for n, v in zip(names, vals):
        a = v.cpu().numpy()
        out[n] = a[:num] if n.startswith(("roi_data", "pool_5", "cls_prob", "bbox_pred")) else a
    out["rois_3d"] = vals[-1]["p3d"][:num].cpu().numpy()
    out["num"] = num
    return out


def oracle_frame_errors(got, params, bv, img, im_info, calib, geom, cfg=None, fv=None):
    """CPU oracle on the same weights/inputs -> dict of measured errors (floats) + exactness flags."""
    from oracle import mv3d_oracle as orc
    from oracle import net_oracle

    keep = {}
    ref = net_oracle.mv3d_test_forward(bv, img, im_info, calib, params, cfg=cfg, geom=geom, keep=keep, fv=fv)
    errs = {}
    for n in FEATURES:
        if n not in got:
            continue
        r = ref[n] if n in ref else keep[n]
        errs[n] = rel(got[n], r)
    if fv is not None:
        errs["conv5_3_3"] = rel(got["conv5_3_3"], ref["conv5_3_3"])
    errs["rpn_cls_prob_abs"] = float(np.abs(got["rpn_cls_prob_reshape"] - ref["rpn_cls_prob_reshape"].numpy()).max())
    # teacher-forced tail: the oracle's pool + head on the GPU's own feature maps and rois
    teacher = dict(conv5_3=got["conv5_3"], conv5_3_2=got["conv5_3_2"],
                   rois=(got["roi_data_bv"], got["roi_data_img"], got["rois_3d"]))
    if fv is not None:
        teacher["conv5_3_3"] = got["conv5_3_3"]
    t = net_oracle.mv3d_test_forward(bv, img, im_info, calib, params, cfg=cfg, geom=geom, fv=fv, teacher=teacher)
    exact = {"pool_5": bool(np.array_equal(got["pool_5"], t["pool_5"])),
             "pool_5_2": bool(np.array_equal(got["pool_5_2"], t["pool_5_2"]))}
    if fv is not None:
        exact["roi_data_fv"] = bool(np.array_equal(got["roi_data_fv"], t["rois_fv"]))
        exact["pool_5_3"] = bool(np.array_equal(got["pool_5_3"], t["pool_5_3"]))
    errs["bbox_pred_tf"] = rel(got["bbox_pred"], t["bbox_pred"])
    errs["cls_prob_tf_abs"] = float(np.abs(got["cls_prob"] - t["cls_prob"].numpy()).max())
    # free-running proposals: the oracle's own RPN outputs -> its own proposals, matched to the GPU's by IoU
    prop = {"oracle_rois": int(ref["rois_bv"].shape[0]), "gpu_rois": int(got["num"]),
            "identical": bool(ref["rois_bv"].shape == got["roi_data_bv"].shape and np.array_equal(ref["rois_bv"], got["roi_data_bv"])),
            "iou90_match": iou_match_fraction(ref["rois_bv"][:, 1:], got["roi_data_bv"][:, 1:], 0.9)}
    # the proposal layer itself on the GPU's RPN outputs (the discrete stage on identical inputs): exact
    with np.errstate(all="ignore"):
        rb, ri, r3 = orc.proposal_layer_3d(got["rpn_cls_prob_reshape"], got["rpn_bbox_pred"], np.asarray(im_info, np.float32),
                                           np.asarray(calib), "TEST", cfg=cfg, geom=geom)
    exact["proposals_on_gpu_rpn_outputs"] = bool(rb.shape == got["roi_data_bv"].shape and np.array_equal(rb, got["roi_data_bv"])
                                                 and np.array_equal(ri, got["roi_data_img"]))
    if not exact["proposals_on_gpu_rpn_outputs"]:   # keep the evidence for an offline look (scratch directory)
        import os
        d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(d, exist_ok=True)
        np.savez_compressed(os.path.join(d, "parity_mismatch_%dviews.npz" % (3 if fv is not None else 2)),
                            prob=got["rpn_cls_prob_reshape"], bbox=got["rpn_bbox_pred"], gpu_bv=got["roi_data_bv"],
                            gpu_img=got["roi_data_img"], gpu_3d=got["rois_3d"], orc_bv=rb, orc_img=ri, orc_3d=r3,
                            im_info=np.asarray(im_info), calib=np.asarray(calib))
    return errs, exact, prop
