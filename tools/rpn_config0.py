"""BASELINE configs[0]: one synthetic KITTI frame, BEV-only 3D-RPN forward (anchors + bbox_transform + projection +
filters + sort + NMS = proposal_layer_3d) -- the reference's algorithm on the host CPU (oracle port, DEVICE=cpu rule:
cpu_nms `>=`) next to the device ProposalLayer3D, same inputs, outputs compared.  Writes gpurun_out/config0_rpn.md.
    python tools/rpn_config0.py [reps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build as ob  # noqa: E402

ob.build()
from oracle import mv3d_oracle as orc  # noqa: E402


def run(reps=2, gpu_iters=20, grids=("REF", "CFG"), keys=("TEST", "TRAIN")):
    """-> list of rows (grid, cfg key, pre, post, proposals, CPU ms, B200 ms, identical)."""
    import torch
    from mv3d_tf_b200.fast_rcnn.config import cfg, cfg_from_end2end_yml
    from mv3d_tf_b200.rpn_msr.proposal_layer_tf import ProposalLayer3D
    from mv3d_tf_b200.utils.transform import CFG_GEOMETRY, REF_GEOMETRY

    cfg_from_end2end_yml()
    cfg.USE_GPU_NMS = False
    gpu = torch.cuda.is_available()
    rows = []
    for (hf, wf), im_info, pg, og, tag in (((75, 75), (601, 601, 1), REF_GEOMETRY, orc.REF_GEOMETRY, "REF 601x601 (N=22500)"),
                                            ((87, 100), (701, 801, 1), CFG_GEOMETRY, orc.CFG_GEOMETRY, "CFG 701x801 (N=34800)")):
        if tag[:3] not in grids:
            continue
        prob, deltas = orc.synth_rpn_outputs(hf, wf, seed=77)
        for key in keys:
            c = cfg[key]
            ocfg = {key: dict(RPN_PRE_NMS_TOP_N=c.RPN_PRE_NMS_TOP_N, RPN_POST_NMS_TOP_N=c.RPN_POST_NMS_TOP_N,
                              RPN_NMS_THRESH=c.RPN_NMS_THRESH, RPN_MIN_SIZE=c.RPN_MIN_SIZE)}
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                bv, img, p3d = orc.proposal_layer_3d(prob, deltas, np.array([im_info], np.float32), orc.KITTI_CALIB, key,
                                                     cfg=ocfg, geom=og, project="loop")   # transform.py:483-500 as the reference runs it
                ts.append(time.perf_counter() - t0)
            cpu_ms = 1e3 * float(np.median(ts))
            gpu_ms, same = float("nan"), "n/a"
            if gpu:
                layer = ProposalLayer3D(hf, wf, key, 8, im_info, geom=pg)
                p, d = torch.from_numpy(prob[0]).cuda(), torch.from_numpy(deltas[0]).cuda()
                for _ in range(3):
                    out = layer(p, d, orc.KITTI_CALIB)
                torch.cuda.synchronize()
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(gpu_iters)]
                for a, b in ev:
                    a.record()
                    out = layer(p, d, orc.KITTI_CALIB)
                    b.record()
                torch.cuda.synchronize()
                gpu_ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
                n = int(out["num"].item())
                same = "yes" if (n == bv.shape[0] and np.array_equal(out["bv"][:n].cpu().numpy(), bv)
                                 and np.array_equal(out["img"][:n].cpu().numpy(), img)) else "NO"
            rows.append((tag, key, c.RPN_PRE_NMS_TOP_N, c.RPN_POST_NMS_TOP_N, bv.shape[0], cpu_ms, gpu_ms, same))
    return rows


def main():
    rows = run(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "config0_rpn.md")
    with open(path, "w") as f:
        f.write("# BASELINE configs[0]: BEV-only 3D-RPN forward (proposal_layer_3d), CPU oracle port (1 thread of %d cores; numpy + C "
                "NMS, the reference's per-box projection loop kept) vs the device layer (CUDA events, median of 20)\n\n" % os.cpu_count())
        f.write("| grid | cfg | pre/post NMS top-N | proposals | CPU ms | B200 ms | survivor set identical |\n|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write("| %s | %s | %d / %d | %d | %.1f | %.3f | %s |\n" % r)
    print(open(path).read())


if __name__ == "__main__":
    main()
