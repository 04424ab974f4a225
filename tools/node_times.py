#!/usr/bin/env python
"""Per-node device time of one MV3D_test frame at the BASELINE shapes (single stream, eager, CUDA events around every
graph node; median over frames) + the raster and, separately, NMS at 6000 -> 300 and 12000 -> 2000 on the frame's own
boxes.  Measurement tool for profiles/ (shares and absolute kernel-group times; not a bench value).
    python tools/node_times.py [--views 2|3] [--mode mixed|precise] [--out profiles/xxx.md]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (synthetic frames + constants)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=2)
    ap.add_argument("--mode", default="mixed")
    ap.add_argument("--frames", type=int, default=7)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from mv3d_tf_b200 import kernels as K
    from mv3d_tf_b200.fast_rcnn.config import cfg, cfg_from_end2end_yml
    from mv3d_tf_b200.networks.factory import get_network
    from mv3d_tf_b200.nms.gpu_nms import nms_device
    from mv3d_tf_b200.utils.read_lidar import BevRasterizer, FvRasterizer
    from mv3d_tf_b200.utils.transform import CFG_GEOMETRY
    from oracle import mv3d_oracle as orc

    cfg_from_end2end_yml()
    cfg.USE_GPU_NMS = False
    net = get_network("MV3D_test", bv_channels=36, precise=True, mixed=(args.mode == "mixed"), geometry=CFG_GEOMETRY,
                      fv=(args.views == 3))
    net.init_weights(seed=7, mode="he")
    net.use_side_stream = False
    raster = BevRasterizer(**bench.BEV)
    fvr = FvRasterizer(net.fv_geometry) if args.views == 3 else None
    im_info = np.array([[701, 801, 1]], np.float32)
    fetch = [net.get_output(n) for n in ("cls_prob", "bbox_pred", "roi_data_bv")]
    frames = [bench.synth_frame(i) for i in range(4)]
    dev = [(torch.from_numpy(p).cuda(), torch.from_numpy(i).cuda()) for p, i in frames]
    fmt = K.FMT_F16E5 if args.mode == "mixed" else K.FMT_BF16X2
    per = {}
    order = []
    raster_ms = []
    for it in range(args.frames + 2):
        pts, img = dev[it % 4]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bv = raster.to_pad(pts, precise=True, fmt=fmt)
        e1.record()
        feed = {net.lidar_bv_data: bv, net.image_data: img, net.im_info: im_info, net.calib: orc.KITTI_CALIB}
        if fvr is not None:
            feed[net.lidar_fv_data] = fvr(pts)[None]
        net.node_events = []
        net.run(fetch, feed)
        torch.cuda.synchronize()
        if it >= 2:
            raster_ms.append(e0.elapsed_time(e1))
            for name, kind, a, b in net.node_events:
                if name not in per:
                    per[name] = (kind, [])
                    order.append(name)
                per[name][1].append(a.elapsed_time(b))
        net.node_events = None
    rows = [("raster (PAD operand format)", "raster", float(np.median(raster_ms)) * 1e3)]
    rows += [(n, per[n][0], float(np.median(per[n][1])) * 1e3) for n in order]
    total = sum(r[2] for r in rows)
    lines = ["# per-node device time, one frame, %d views, %s mode (eager, one stream, CUDA events incl. launch gaps; median of %d frames)"
             % (args.views, args.mode, args.frames), "", "| node | kind | us | share |", "|---|---|---|---|"]
    for n, k, us in rows:
        lines.append("| %s | %s | %.1f | %.1f%% |" % (n, k, us, 100 * us / total))
    lines.append("| **sum** | | %.1f | |" % total)
    by_kind = {}
    for n, k, us in rows:
        by_kind[k] = by_kind.get(k, 0.0) + us
    lines += ["", "| kind | us | share |", "|---|---|---|"] + ["| %s | %.1f | %.1f%% |" % (k, v, 100 * v / total)
                                                                for k, v in sorted(by_kind.items(), key=lambda kv: -kv[1])]
    # NMS alone on realistic boxes: the frame's decoded proposals, sorted by score
    from mv3d_tf_b200.rpn_msr.proposal_layer_tf import ProposalLayer3D
    prob, deltas = orc.synth_rpn_outputs(87, 100, seed=77)
    lines += ["", "| NMS on %s | us (median of 20, CUDA events) | kept |" % "synthetic RPN outputs (87x100x4 anchors)", "|---|---|---|"]
    for key, pre, post in (("TEST", 6000, 300), ("TRAIN", 12000, 2000)):
        layer = ProposalLayer3D(87, 100, key, 8, (701, 801, 1), geom=CFG_GEOMETRY)
        st = layer.decode(torch.from_numpy(prob[0]).cuda(), torch.from_numpy(deltas[0]).cuda(), orc.KITTI_CALIB)
        keep = st["keep"].bool()
        sc = st["score"][keep]
        boxes = st["pbv"][keep][torch.argsort(sc, descending=True)][:pre].contiguous()
        ts = []
        for _ in range(23):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            k, num = nms_device(boxes, 0.7, True, max_keep=post)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        lines.append("| %d boxes -> %d | %.1f | %d |" % (boxes.shape[0], post, float(np.median(ts[3:])), int(num.item())))
    text = "\n".join(lines) + "\n"
    print(text)
    if args.out:
        open(args.out, "w").write(text)


if __name__ == "__main__":
    main()
