"""Per-layer timing of the conv/fc GEMMs of one BASELINE-shaped frame (CUDA events on the launch stream,
median of 5 frames after 3 warm-ups).  Writes profiles/layers_<tag>.md.   python tools/profile_layers.py [tag] [mode]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mv3d_tf_b200 import kernels  # noqa: E402
from mv3d_tf_b200.fast_rcnn.config import cfg, cfg_from_end2end_yml  # noqa: E402
from mv3d_tf_b200.networks.factory import get_network  # noqa: E402
from mv3d_tf_b200.utils.read_lidar import BevRasterizer  # noqa: E402
from mv3d_tf_b200.utils.transform import CFG_GEOMETRY  # noqa: E402
from oracle import mv3d_oracle as orc  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    mode = sys.argv[2] if len(sys.argv) > 2 else "precise"
    cfg_from_end2end_yml()
    cfg.USE_GPU_NMS = False
    net = get_network("MV3D_test", bv_channels=36, precise=(mode != "fast"), mixed=(mode == "mixed"), geometry=CFG_GEOMETRY)
    net.init_weights(seed=7, mode="he")
    raster = BevRasterizer(**bench.BEV)
    pts, img = bench.synth_frame(0)
    pts, img = torch.from_numpy(pts).cuda(), torch.from_numpy(img).cuda()
    im_info = np.array([[701, 801, 1]], np.float32)
    fetch = [net.get_output("cls_prob"), net.get_output("bbox_pred")]
    log = []
    orig = kernels._run_gemm

    def traced(**kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        orig(**kw)
        e1.record(torch.cuda.current_stream())
        log.append((kw["M"], kw["N"], kw["Cin"] * kw["taps"], kw["taps"], kw["split_k"], e0, e1))
    reps = 8
    for i in range(reps):
        if i == 3:
            kernels._run_gemm = traced
        bv = raster.to_pad(pts, precise=(mode != "fast"))
        net.run(fetch, {net.lidar_bv_data: bv, net.image_data: img, net.im_info: im_info, net.calib: orc.KITTI_CALIB})
    torch.cuda.synchronize()
    kernels._run_gemm = orig
    n = len(log) // (reps - 3)
    names = [nd.name for nd in net._program if nd.kind in ("conv", "fc") and "fused_into" not in nd.attrs]
    rows, tot_ms, tot_fl = [], 0.0, 0.0
    for j in range(n):
        M, N, Kd, taps, split = log[j][:5]
        ms = float(np.median([log[j + r * n][5].elapsed_time(log[j + r * n][6]) for r in range(reps - 3)]))
        fl = 2.0 * M * N * Kd
        rows.append((names[j] if j < len(names) else "?", M, N, Kd, split, ms, fl / ms / 1e9))
        tot_ms += ms
        tot_fl += fl
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = os.path.join(ROOT, "gpurun_out", "layers_%s_%s.md" % (tag, mode))
    with open(out, "w") as f:
        f.write("# per-layer conv/fc GEMM timing, mode=%s (CUDA events, median of 5 frames; TFLOP/s counts padded "
                "M,K once -- multiply by 3 for the MMA rate in precise mode)\n\n" % mode)
        f.write("| layer | M | N | K | split_k | ms | TFLOP/s (1x) |\n|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write("| %s | %d | %d | %d | %d | %.4f | %.1f |\n" % r)
        f.write("| **total** | | | | | %.3f | %.1f |\n" % (tot_ms, tot_fl / tot_ms / 1e9))
    print(open(out).read())


if __name__ == "__main__":
    main()
