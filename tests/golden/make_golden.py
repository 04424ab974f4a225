"""Generate tests/golden/*.npz by running the REFERENCE's own sources (through oracle/ref_shim.py:
mechanical py2->py3 patches, arithmetic untouched) on seeded synthetic inputs.

Run in the dev container only (needs /root/reference):   python tests/golden/make_golden.py
The fixtures are committed; the GPU box and CI never need the reference tree.
Inputs are stored next to outputs so the fixtures do not depend on RNG stream stability.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import mv3d_oracle as orc  # noqa: E402  (synthetic input generators only)
from oracle import ref_shim  # noqa: E402


def main():
    R = ref_shim.load()
    R.config.cfg_from_file(R.yml)  # the live overlay: experiments/cfgs/faster_rcnn_end2end.yml
    out = {}

    # ---- raster: reference point_cloud_2_top on two small grids (zres 0.3 / 0.1) ----------
    pts = orc.synth_points(30000, seed=11)
    pts[:, 0] = pts[:, 0] * 0.25  # concentrate inside the small crop: many duplicate cells
    pts[:, 1] = pts[:, 1] * 0.25
    # exact boundary probes: z on slice bounds, x/y on crop bounds, cell edges
    edge = np.array([[5.0, 0.0, -2.0, .5], [5.0, 1.0, -1.7, .6], [5.0, 2.0, 0.4, .7], [0.0, 0.0, 0, .1],
                     [16.0, 0.0, 0, .2], [3.0, 8.0, 0, .3], [3.0, -8.0, 0, .4], [0.05, 0.05, -0.5, .8],
                     [0.05, -0.05, -0.5, .9], [15.95, 7.95, 0.39999998, .11], [1.0, 1.0, 0.1, .12],
                     [1.0, 1.0, 0.099999994, .13], [2.3, -4.1, -1.1, .14], [2.3, -4.1, -1.1, .15]], dtype=np.float32)
    pts = np.vstack((pts, edge)).astype(np.float32)
    ra = dict(res=0.1, zres=0.3, side_range=(-8., 8.), fwd_range=(0., 16.), height_range=(-2, 0.4))
    rb = dict(res=0.1, zres=0.1, side_range=(-8., 8.), fwd_range=(0., 12.), height_range=(-2.0, 1.5))
    top_a = R.read_lidar.point_cloud_2_top(pts, **ra)
    top_b = R.read_lidar.point_cloud_2_top(pts, **rb)
    np.savez_compressed(os.path.join(HERE, "raster.npz"), points=pts, top_a=top_a, top_b=top_b)
    print("raster", top_a.shape, top_b.shape, int((top_a != 0).sum()), int((top_b != 0).sum()))

    # ---- anchors ---------------------------------------------------------------------------
    base = R.generate_anchors.generate_anchors_bv()

    # ---- proposal chain on a 40x44 feature map (reference grid constants: 0..60, -30..30) --
    hf, wf = 40, 44
    prob, deltas = orc.synth_rpn_outputs(hf, wf, seed=21)
    # exercise overflow / degenerate paths (SURVEY A11 ii, iii): a few large deltas
    d = deltas.reshape(-1, 6)
    d[5, 3] = 30.0
    d[6, 0] = -40.0
    d[7, 4] = 100.0  # exp overflow -> inf
    im_info = np.array([[601, 601, 1]], dtype=np.float32)
    calib = orc.KITTI_CALIB
    T, B = R.transform, R.bbox_transform
    anchors = orc.enumerate_anchors(hf, wf, 8, base)
    with np.errstate(all="ignore"):
        a3d = T.bv_anchor_to_lidar(anchors)
        p3d = B.bbox_transform_inv_3d(a3d, deltas.reshape(-1, 6))
        pbv = T.lidar_3d_to_bv(p3d)
        cnr = T.lidar_3d_to_corners(p3d)
        pimg = T.lidar_cnr_to_img(cnr, calib[3], calib[2], calib[0])
        blobs = {}
        for key in ("TEST", "TRAIN"):
            blobs[key] = R.proposal_layer_tf.proposal_layer_3d(prob, deltas, im_info, calib, key, [8, ], [1.0, 1.0])
    np.savez_compressed(
        os.path.join(HERE, "proposal.npz"), base_anchors=base, prob=prob, deltas=deltas, im_info=im_info, calib=calib,
        anchors_3d=a3d, p3d=p3d, pbv=pbv, corners=cnr, pimg=pimg,
        test_bv=blobs["TEST"][0], test_img=blobs["TEST"][1], test_3d=blobs["TEST"][2],
        train_bv=blobs["TRAIN"][0], train_img=blobs["TRAIN"][1], train_3d=blobs["TRAIN"][2],
        cfg=np.array([R.cfg.TEST.RPN_PRE_NMS_TOP_N, R.cfg.TEST.RPN_POST_NMS_TOP_N, R.cfg.TRAIN.RPN_PRE_NMS_TOP_N,
                      R.cfg.TRAIN.RPN_POST_NMS_TOP_N]))
    print("proposal", [b.shape for b in blobs["TEST"]], [b.shape for b in blobs["TRAIN"]])

    # ---- NMS (cpu_nms.pyx, utils/nms.pyx) and IoU (bbox.pyx) -------------------------------
    rng = np.random.default_rng(31)
    n = 1500
    x1 = rng.integers(0, 560, n).astype(np.float32)
    y1 = rng.integers(0, 560, n).astype(np.float32)
    w = rng.integers(4, 60, n).astype(np.float32)
    h = rng.integers(4, 60, n).astype(np.float32)
    sc = rng.permutation(n).astype(np.float32) / n  # tie-free
    dets = np.stack((x1, y1, np.minimum(x1 + w, 600), np.minimum(y1 + h, 600), sc), 1).astype(np.float32)
    dets[100:140, :4] = dets[60:100, :4]  # exact duplicates -> IoU == 1
    keep07 = np.asarray(R.cpu_nms.cpu_nms(dets, 0.7), dtype=np.int32)
    keep05 = np.asarray(R.cpu_nms.cpu_nms(dets, 0.5), dtype=np.int32)
    keep01 = np.asarray(R.cython_nms.nms(dets, 0.1), dtype=np.int32)
    keep_new = np.asarray(R.cython_nms.nms_new(dets, 0.3), dtype=np.int32)
    boxes = dets[:400, :4].astype(np.float64)
    query = dets[400:420, :4].astype(np.float64)
    iou = R.cython_bbox.bbox_overlaps(np.ascontiguousarray(boxes), np.ascontiguousarray(query))
    np.savez_compressed(os.path.join(HERE, "nms_iou.npz"), dets=dets, keep07=keep07, keep05=keep05, keep01=keep01,
                        keep_new=keep_new, boxes=boxes, query=query, iou=iou)
    print("nms", len(keep07), len(keep05), len(keep01), len(keep_new))


if __name__ == "__main__":
    main()
