#!/usr/bin/env python
"""BASELINE configs[4]: LiDAR -> BEV rasterisation throughput sweep, 10k-1M points/frame on the 701x801x36 grid.

    python tools/raster_sweep.py [--out profiles/xxx.md] [--gpus N]     (N>1: launch under torchrun; frames are independent,
                                                                      every rank rasterises its own clouds, no collective)
Achieved GB/s = algorithmic bytes (16 B/point read once + 4*H*W*C output bytes written once) / CUDA-event time,
against the measured HBM copy peak in MEASURED_PEAKS.json.  Four output buffers rotate (324 MB > the 126 MB L2) so that
write-backs are paid inside the timed region.  Every size is first checked bit-exact against the CPU oracle.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BEV = dict(res=0.1, zres=0.1, side_range=(-40., 40.), fwd_range=(0., 70.), height_range=(-2.0, 1.5))


def sweep(sizes=(10000, 30000, 100000, 120000, 300000, 1000000), iters=40, check=True, rank=0, world=1):
    """-> list of row dicts; under torch.distributed (world > 1) the time is the max over ranks."""
    from mv3d_tf_b200.utils.read_lidar import BevRasterizer
    from oracle import build as ob
    ob.build()
    from oracle import mv3d_oracle as orc   # checker + synthetic clouds only

    peak = 6536.4
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    r = BevRasterizer(**BEV)
    H, W, C = r.shape
    outs = [torch.empty(r.shape, dtype=torch.float32, device="cuda") for _ in range(4)]
    rows = []
    for n in sizes:
        pts_h = orc.synth_points(n, seed=77 + rank)
        pts = torch.from_numpy(pts_h).cuda()
        if check and n <= 300000:
            assert np.array_equal(r(pts).cpu().numpy(), orc.point_cloud_2_top(pts_h, **BEV)), "raster mismatch at n=%d" % n
        for i in range(5):
            r(pts, out=outs[i % 4])
        torch.cuda.synchronize()
        # the frames are replayed from one CUDA graph (as FrameRunner does in production) so that the small launches
        # per frame are not paced by Python/ctypes launch latency
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for i in range(iters):
                r(pts, out=outs[i % 4])
        torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        algo = 16.0 * n + 4.0 * H * W * C
        gbs = algo / (ms * 1e-3) / 1e9
        rows.append(dict(points=n, ms=ms, algorithmic_MB=algo / 1e6, GBps=gbs, frac_of_measured_hbm=gbs / peak,
                         frames_per_s=world * 1e3 / ms))
    return rows, (H, W, C), peak


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")
    rows, (H, W, C), peak = sweep(iters=args.iters, check=not args.no_check, rank=rank, world=world)
    if rank == 0:
        lines = ["# LiDAR -> BEV raster sweep (configs[4]), %d GPU(s), grid %dx%dx%d, float32 (H,W,C) output" % (world, H, W, C), "",
                 "time = CUDA events around one CUDA-graph replay of %d frames, max over ranks; peak = %.1f GB/s (MEASURED_PEAKS.json hbm_gbs)" % (args.iters, peak),
                 "", "| points | ms/frame | algorithmic MB | achieved GB/s | of measured HBM | frames/s (all GPUs) |", "|---|---|---|---|---|---|"]
        for x in rows:
            lines.append("| %d | %.4f | %.1f | %.0f | %.3f | %.0f |" % (x["points"], x["ms"], x["algorithmic_MB"], x["GBps"],
                                                                          x["frac_of_measured_hbm"], x["frames_per_s"]))
        text = "\n".join(lines) + "\n"
        print(text)
        print(json.dumps({"raster_sweep": rows, "n_gpus": world}))
        if args.out:
            with open(args.out, "w") as f:
                f.write(text)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
