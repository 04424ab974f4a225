#!/usr/bin/env python
"""MV3D per-frame detection hot path: frames/s on synthetic KITTI-shaped frames (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode mixed|precise|fast]

One step = one frame through the whole path: LiDAR (120 k points) -> BEV raster 701x801x36 -> BEV + RGB
VGG16 trunks (tcgen05 implicit GEMM) -> RPN head -> device proposal layer (decode/sort/NMS) -> fused 2-view
ROI pool -> fc6/fc7 x2 -> cls_prob / bbox_pred.   N>1: frames shard data-parallel, one process per GPU, no
collective in inference ("weak" scaling: K frames per rank).
Prints ONE JSON line (see the task contract): value (inputs resident in HBM), e2e (host buffers in, results
out), roofline of the dominant kernel (the conv/fc GEMM), cpu_baseline (the oracle port on the host cores).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BEV = dict(res=0.1, zres=0.1, side_range=(-40., 40.), fwd_range=(0., 70.), height_range=(-2.0, 1.5))  # 701x801x36
IMG_HW = (375, 1242)
N_POINTS = 120000
PIXEL_MEANS = np.array([95.8814, 98.7743, 93.8549], np.float32)


def synth_frame(frame_id):
    from oracle import mv3d_oracle as orc  # synthetic-input generator shared with the tests (not compute)

    pts = orc.synth_points(N_POINTS, seed=1234 + frame_id)
    rng = np.random.default_rng(99 + frame_id)
    img = rng.integers(0, 256, (1, IMG_HW[0], IMG_HW[1], 3)).astype(np.float32) - PIXEL_MEANS
    return pts, img.astype(np.float32)


def gemm_flops(net, hb, wb, hi, wi, R):
    """Algorithmic FLOPs of every conv / fc GEMM of one frame: sum 2*k*k*Cin*Cout*Hout*Wout."""
    total = 0

    def trunk(h, w, cin):
        t = 0
        chans = [64, 64, 'p', 128, 128, 'p', 256, 256, 256, 'p', 512, 512, 512, 512, 512, 512]
        c = cin
        for x in chans:
            if x == 'p':
                h, w = h // 2, w // 2
            else:
                t += 2 * 9 * c * x * h * w
                c = x
        return t, h, w
    tb, hf, wf = trunk(hb, wb, net.lidar_bv_data.channels)
    ti, _, _ = trunk(hi, wi, 3)
    total += tb + ti
    views = 2
    if getattr(net, "with_fv", False):
        total += trunk(net.fv_geometry.H, net.fv_geometry.W, 3)[0]
        views = 3
    total += 2 * 9 * 512 * 512 * hf * wf + 2 * 512 * (8 + 24) * hf * wf          # RPN head
    total += views * (2 * R * 25088 * 2048 + 2 * R * 2048 * 2048) + 2 * R * (2048 * views) * 50  # fusion head
    return total


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling guide recipe): one nvidia-smi
    process in loop mode (-lms 50), its lines collected by a reader thread."""

    def __init__(self, index):
        self.rows, self.index, self.proc = [], index, None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.strip().split(",")])
        except Exception:
            pass

    def __enter__(self):
        self.th.start()
        time.sleep(0.15)   # let the first sample land before the timed region starts
        return self

    def __exit__(self, *a):
        try:
            if self.proc is not None:
                self.proc.terminate()
        except Exception:
            pass
        self.th.join(timeout=3)

    def summary(self):
        sm = [int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        pw = []
        for r in self.rows:
            try:
                pw.append(float(r[6]))
            except Exception:
                pass
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ---------------------------------------------------------------------------------------------------------
def cpu_frame(orc, net_oracle, params, pts, img, im_info, calib, cfg, views=2):
    """One frame of the reference's CPU path: the oracle port (numpy/C restatement + torch-CPU fp32 graph),
    with the per-box projection loop in the reference's own shape (transform.py:483-500)."""
    bv = orc.point_cloud_2_top(pts, **BEV)[None]
    import torch

    with torch.no_grad():
        c5 = net_oracle.trunk(bv, params, "")
        c5_2 = net_oracle.trunk(img, params, "_2")
        prob, bbox = net_oracle.rpn_head(c5, params)
        rois_bv, rois_img, rois_3d = orc.proposal_layer_3d(prob.numpy(), bbox.numpy(), im_info, calib, "TEST", cfg=cfg,
                                                           geom=orc.CFG_GEOMETRY, project="loop")
        p1, _ = orc.roi_pool_fwd(c5.numpy(), rois_bv)
        p2, _ = orc.roi_pool_fwd(c5_2.numpy(), rois_img)
        p3 = None
        if views == 3:   # front view: our specification (the reference has none), same oracle port
            c5_3 = net_oracle.trunk(orc.point_cloud_2_front(pts)[None], params, "_3")
            rois_fv = np.hstack((rois_3d[:, :1], orc.lidar_3d_to_fv(rois_3d[:, 1:7]))).astype(np.float32)
            p3, _ = orc.roi_pool_fwd(c5_3.numpy(), rois_fv)
        cls, bb = net_oracle.fusion_head(p1, p2, params, pool_fv=p3)
    return cls, bb


def make_cpu_params(seed=7, views=2):
    """Same architecture, random init, built directly on the host (no GPU needed for the reference arm)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    params = {}

    def add(name, shape):
        fan_in = int(np.prod(shape[:-1]))
        w = torch.empty(shape)
        torch.nn.init.trunc_normal_(w, 0.0, 1.0, -2.0, 2.0, generator=g)
        params[name] = dict(weights=(w * (2.0 / fan_in) ** 0.5).numpy(), biases=np.zeros(shape[-1], np.float32))
    for suffix, cin in (("", 36), ("_2", 3), ("_3", 3))[:views]:
        c = cin
        for item in ("conv1_1", 64), ("conv1_2", 64), ("conv2_1", 128), ("conv2_2", 128), ("conv3_1", 256), \
                ("conv3_2", 256), ("conv3_3", 256), ("conv4_1", 512), ("conv4_2", 512), ("conv4_3", 512), \
                ("conv5_1", 512), ("conv5_2", 512), ("conv5_3", 512):
            add(item[0] + suffix, (3, 3, c, item[1]))
            c = item[1]
    add("rpn_conv/3x3", (3, 3, 512, 512)); add("rpn_cls_score", (1, 1, 512, 8)); add("rpn_bbox_pred", (1, 1, 512, 24))
    for b in ("_1", "_2", "_3")[:views]:
        add("fc6" + b, (25088, 2048)); add("fc7" + b, (2048, 2048))
    add("cls_score", (2048 * views, 2)); add("bbox_pred", (2048 * views, 48))
    return params


def run_cpu_reference(steps, warmup, frames, views=2, params=None):
    import torch

    from oracle import build as ob
    ob.build()
    from oracle import mv3d_oracle as orc
    from oracle import net_oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = make_cpu_params(views=views) if params is None else params
    cfg = {"TEST": dict(RPN_PRE_NMS_TOP_N=6000, RPN_POST_NMS_TOP_N=300, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=5)}
    im_info = np.array([[701, 801, 1]], np.float32)
    t_w = time.perf_counter()
    for i in range(max(1, warmup)):
        cpu_frame(orc, net_oracle, params, *frames[i % len(frames)], im_info, orc.KITTI_CALIB, cfg, views)
    t_frame = (time.perf_counter() - t_w) / max(1, warmup)
    # bounded sample: whole frames, as many of the requested steps as fit in ~150 s of CPU work (at least 2)
    steps = max(2, min(steps, int(150.0 / max(t_frame, 1e-3))))
    t0 = time.perf_counter()
    for i in range(steps):
        cpu_frame(orc, net_oracle, params, *frames[i % len(frames)], im_info, orc.KITTI_CALIB, cfg, views)
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps * 1e3, cores, steps


def make_config(views):
    """The `config` object of the JSON line -- identical in both arms (ours / reference) for a given --views."""
    net = ("the reference's own two-view network (BEV + RGB; lib/networks/network.py:313-315 has no front-view branch)"
           if views == 2 else "BEV + FV (64x512x3, this repo's written spec: the reference has none) + RGB")
    return {"workload": "configs[1]: full MV3D inference batch=1 per GPU, %s: 120k-pt LiDAR -> BEV 701x801x36 raster, "
                        "VGG16 trunks (RGB 375x1242), 3D-RPN + proposal layer (6000/300, NMS 0.7), fused multi-view ROI "
                        "pool, fc fusion head" % net,
            "views": views, "frames_per_step_per_gpu": 1, "parallelism": "frames data-parallel, no collective",
            "l2_policy": "per-frame working set (144 MB/activation at conv1, 411 MB fc6 weights) exceeds the 126 MB L2; "
                         "4 distinct frames rotate"}


def run_train_bench(args, rank, world, local_rank):
    """BASELINE configs[2]: MV3D train step fwd+bwd+Adam, batch = 2 frames per GPU (N>1: one NCCL all-reduce of the flat
    gradient buffer per step).  Returns the dict that goes under "train_step" (and is the main line with --workload train)."""
    import torch
    import torch.distributed as dist

    from mv3d_tf_b200 import _lib, kernels
    from mv3d_tf_b200.fast_rcnn.config import cfg, cfg_from_end2end_yml
    from mv3d_tf_b200.fast_rcnn.train_mv import SolverWrapper
    from mv3d_tf_b200.networks.factory import get_network
    from mv3d_tf_b200.utils.read_lidar import BevRasterizer
    from mv3d_tf_b200.utils.transform import CFG_GEOMETRY
    from oracle import mv3d_oracle as orc  # synthetic GT generator + calib constants only

    cfg_from_end2end_yml()
    cfg.USE_GPU_NMS = False
    B = args.train_batch
    precise = args.mode != "fast"   # training keeps the bf16 hi/lo 3-pass GEMMs in both parity modes
    net = get_network("MV3D_train", bv_channels=36, precise=precise, geometry=CFG_GEOMETRY)
    net.init_weights(seed=7, mode="he")
    sw = SolverWrapper(network=net, keep_prob=0.5, process_group=dist.group.WORLD if world > 1 else None)
    raster = BevRasterizer(**BEV)
    n_sets = 2
    sets = []
    for sidx in range(n_sets):
        fr = [synth_frame(100 + 16 * rank + sidx * B + b) for b in range(B)]
        gts = [orc.synth_gt(6, seed=500 + 16 * rank + sidx * B + b, geom=orc.CFG_GEOMETRY) for b in range(B)]
        sets.append(dict(pts=[torch.from_numpy(f[0]).cuda() for f in fr],
                         img=torch.from_numpy(np.concatenate([f[1] for f in fr])).cuda(),
                         gt_bv=[g[0] for g in gts], gt_3d=[g[1] for g in gts], gt_cnr=[g[2] for g in gts]))
    im_info = np.array([[701, 801, 1]], np.float32)
    np.random.seed(3 + rank)   # tools/train_net.py:78-80 seeds numpy with cfg.RNG_SEED = 3

    def step(i):
        s_ = sets[i % n_sets]
        bv = raster.to_pad(s_["pts"], precise=precise)
        return sw.train_step(dict(image_data=s_["img"], lidar_bv_data=bv, im_info=im_info, gt_boxes_bv=s_["gt_bv"],
                                  gt_boxes_3d=s_["gt_3d"], gt_boxes_corners=s_["gt_cnr"], calib=orc.KITTI_CALIB))

    stream = torch.cuda.current_stream()
    for i in range(max(args.warmup, 3)):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import gc
    gc.collect()
    gc.disable()     # a generation-2 collection inside the 10 timed steps shows up as one 30+ ms step (Python, not the GPU)
    e0.record(stream)
    t_host = time.perf_counter()
    per_step = []
    for i in range(args.train_steps):
        loss = step(i)
        ev = torch.cuda.Event(enable_timing=True)   # per-step device times (one event per step)
        ev.record(stream)
        per_step.append(ev)
    host_ms = (time.perf_counter() - t_host) * 1e3 / args.train_steps   # Python + host sampling + its two D2H syncs per step
    e1.record(stream)
    torch.cuda.synchronize()
    gc.enable()
    ms = e0.elapsed_time(e1) / args.train_steps
    ts = [e0.elapsed_time(ev) for ev in per_step]
    step_ms = [b - a for a, b in zip([0.0] + ts[:-1], ts)]
    if os.environ.get("MV3D_TRAIN_TRACE"):
        print("train per-step ms: median %.2f | %s" % (float(np.median(step_ms)), " ".join("%.1f" % x for x in step_ms)), file=sys.stderr)
    launches = _lib.launch_count() // args.train_steps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    # per-GEMM-launch events: forward GEMMs, backward-data GEMMs, backward-filter GEMMs
    kernels.GEMM_EVENTS = []
    step(0)
    torch.cuda.synchronize()
    gemm_ms = sum(ev[0].elapsed_time(ev[1]) for ev in kernels.GEMM_EVENTS)
    n_gemm = len(kernels.GEMM_EVENTS)
    per_kernel = {}
    for ev in kernels.GEMM_EVENTS:
        k = per_kernel.setdefault(ev[2], [0, 0.0, 0.0])
        k[0] += 1; k[1] += ev[0].elapsed_time(ev[1]); k[2] += ev[3]
    kernels.GEMM_EVENTS = None
    gemm_by_kernel = {k: {"launches": v[0], "ms": v[1], "tflops_1x": v[2] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0}
                      for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][1])}
    return {"ms": ms, "ms_median_step": float(np.median(step_ms)), "gemm_by_kernel": gemm_by_kernel, "frames_per_step_per_gpu": B, "frames_per_s": world * B / (ms * 1e-3), "steps": args.train_steps,
            "n_gpus": world, "gpu_launches_per_step": launches, "host_enqueue_ms_per_step": host_ms,
            "gemm_ms": gemm_ms, "gemm_launches": n_gemm,
            "loss": [float(x) for x in loss.tolist()], "optimizer": "Adam lr=1e-5 (TF-1.0 defaults), keep_prob 0.5",
            "grad_allreduce": "NCCL all-reduce of the flat fp32 gradient buffer (%.0f MB)" % (sw.grad.numel() * 4 / 1e6)
            if world > 1 else "off (single GPU)", "mode": "precise" if args.mode != "fast" else "fast",
            "workload": "configs[2]: MV3D train step fwd+bwd+Adam, %d frames/GPU: 120k-pt LiDAR -> BEV 701x801x36, RGB "
                        "375x1242, 6 GT cars/frame, RPN 12000/2000 proposals, 128 sampled rois/frame" % B}


def cpu_train_step_ms(frames, params, views_unused=2):
    """The reference's train step on the host cores (oracle port: torch-CPU fp32 autograd + numpy/C target layers),
    ONE frame timed; a configs[2] step is two such frames back to back (the reference is strictly batch 1)."""
    import torch
    from oracle import mv3d_oracle as orc
    from oracle import net_oracle

    torch.set_num_threads(os.cpu_count() or 1)
    pts, img = frames[0]
    bv = orc.point_cloud_2_top(pts, **BEV)[None]
    gt = orc.synth_gt(6, seed=500, geom=orc.CFG_GEOMETRY)
    np.random.seed(3)
    t0 = time.perf_counter()
    with np.errstate(all="ignore"):
        net_oracle.train_forward_backward(bv, img, np.array([[701, 801, 1]], np.float32), orc.KITTI_CALIB, *gt, params,
                                          geom=orc.CFG_GEOMETRY)
    return (time.perf_counter() - t0) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="mixed", choices=["precise", "mixed", "fast"],
                    help="precise: bf16 hi/lo 3-pass everywhere; mixed: fp16 + e5m2-pair operands (2 pass-equivalents) "
                         "in the 3x3 convs, 3-pass elsewhere; fast: single bf16 pass (not a parity mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="infer", choices=["infer", "train"],
                    help="infer: configs[1] frames/s (the headline); train: configs[2] train-step ms only")
    ap.add_argument("--train-steps", type=int, default=10, help="timed train steps reported under train_step")
    ap.add_argument("--train-batch", type=int, default=2)
    ap.add_argument("--no-train", action="store_true", help="skip the train_step leg of the default run")
    ap.add_argument("--no-precise-leg", action="store_true",
                    help="mixed mode only: skip the extra device-resident timing of the bf16x3-everywhere network")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[0] / configs[4] extras")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured CUDA graph")
    ap.add_argument("--in-flight", type=int, default=2, help="independent batch-1 frames in flight per GPU (streams)")
    ap.add_argument("--views", type=int, default=2, choices=[2, 3],
                    help="2 (default): the reference's own BEV+RGB network (lib/networks/network.py:313-315 has no FV "
                         "branch) -- the anchored comparison; 3: BEV+FV+RGB.  The other view count is reported beside it")
    ap.add_argument("--no-other-views", action="store_true", help="skip the extra leg with the other view count")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)
    n_frames = 4
    frames = [synth_frame(i + 16 * rank) for i in range(n_frames)]
    other_views = 5 - args.views

    if args.impl == "reference":
        if rank != 0:
            return
        fps, ms, cores, n_done = run_cpu_reference(args.steps, warmup, frames, args.views)
        line = {"impl": "reference", "metric": "MV3D inference frames/sec", "value": fps, "unit": "frames/s",
                "n_gpus": args.gpus, "steps": n_done, "warmup": warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": make_config(args.views),
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                 "sample": "%d whole frames, 1 per step (of %d requested; bounded to ~150 s; numpy/C "
                                           "oracle port of the reference's host layers incl. its per-box projection "
                                           "loop; conv/fc via torch-CPU fp32 because TensorFlow 1.0 is not installable); "
                                           "one CPU process whatever --gpus says" % (n_done, args.steps)},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if not args.no_other_views:
            f2, m2, _, n2 = run_cpu_reference(2, 1, frames, other_views)
            line["views%d" % other_views] = {"value": f2, "unit": "frames/s", "steps": n2, "config": make_config(other_views)}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from mv3d_tf_b200 import _lib, kernels
    from mv3d_tf_b200.fast_rcnn.config import cfg, cfg_from_end2end_yml
    from mv3d_tf_b200.networks.factory import get_network
    from mv3d_tf_b200.utils.read_lidar import BevRasterizer, FvRasterizer
    from mv3d_tf_b200.utils.transform import CFG_GEOMETRY
    from oracle import mv3d_oracle as orc  # KITTI_CALIB constants + synthetic generators only

    if args.workload == "train":
        tr = run_train_bench(args, rank, world, local_rank)
        if rank == 0:
            print(json.dumps({"metric": "MV3D train-step ms", "value": tr["ms"], "unit": "ms", "n_gpus": world,
                              "steps": tr["steps"], "warmup": max(args.warmup, 3), "ms_per_step": tr["ms"],
                              "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "bf16x3",
                              "data": "synthetic", "config": {"workload": tr["workload"], "mode": "precise" if args.mode != "fast" else "fast"},
                              "gpu_launches": tr["gpu_launches_per_step"] * tr["steps"], "train_step": tr}))
        if world > 1:
            dist.destroy_process_group()
        return

    cfg_from_end2end_yml()
    cfg.USE_GPU_NMS = False  # reproduce the DEVICE=cpu rule (cpu_nms `>=`), the parity target
    im_info = np.array([[701, 801, 1]], np.float32)
    calib = orc.KITTI_CALIB
    stream = torch.cuda.current_stream()
    dev_frames = [(torch.from_numpy(p).cuda(), torch.from_numpy(i).cuda()) for p, i in frames]
    pin_frames = [(torch.from_numpy(p).pin_memory(), torch.from_numpy(i).pin_memory()) for p, i in frames]
    depth = max(1, args.in_flight)
    want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline

    from mv3d_tf_b200.fast_rcnn.test_mv import FramePipeline, FrameRunner

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class Leg:
        """One network (views, mode) behind the product call: each frame is one captured CUDA graph (FrameRunner);
        `--in-flight` of them are replayed round-robin on separate streams (FramePipeline) so that one frame's serial
        tail overlaps the next frame's trunks."""

        def __init__(self, views, mode):
            self.views, self.mode = views, mode
            self.net = get_network("MV3D_test", bv_channels=36, precise=(mode != "fast"), mixed=(mode == "mixed"),
                                   geometry=CFG_GEOMETRY, fv=(views == 3))
            self.net.init_weights(seed=7, mode="he")
            _lib.reset_launch_count()
            self.runner = self.make_runner().capture()
            self.launches_per_frame = _lib.launch_count() // (3 if not args.no_graph else 2)
            self.pipe = FramePipeline(self.make_runner, depth=depth)

        def make_runner(self):
            r = FrameRunner(self.net, BevRasterizer(**BEV), N_POINTS, IMG_HW, im_info,
                            fetch=("cls_prob", "bbox_pred", "roi_data_bv"), use_graph=not args.no_graph)
            r.load_device(dev_frames[0][0], dev_frames[0][1], calib)
            return r

        def step_device(self, i):
            pts, img = dev_frames[i % n_frames]
            pipe = self.pipe
            if pipe.head - pipe.tail >= depth:
                pipe.tail += 1                      # device-resident leg: nothing to collect on the host
            pipe.submit(pts, img, None, device_inputs=True)   # D2D into the graph's static inputs (4 distinct frames rotate)

        def step_e2e(self, i):
            pts_h, img_h = pin_frames[i % n_frames]
            pipe = self.pipe
            if pipe.head - pipe.tail >= depth:
                pipe.collect()                      # the oldest frame's detections are on the host (pinned) here
            pipe.submit(pts_h, img_h, calib)        # pinned host in -> H2D, graph, D2H -> pinned host out

        def timed(self, fn, steps):
            pipe = self.pipe
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream)
            for i in range(steps):
                fn(i)
            while pipe.tail < pipe.head and fn == self.step_e2e:
                pipe.collect()                      # the last frames' detections reach the host inside the timed region
            pipe.drain()                            # device-side join of the per-frame streams
            e1.record(stream)
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            t = torch.tensor([max(e0.elapsed_time(e1), 0.0), wall * 1e3], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0]), float(t[1])

        def measure(self, steps, e2e=True):
            for i in range(warmup):
                self.step_device(i)
            ms_dev, _ = self.timed(self.step_device, steps)
            out = {"value": world * steps / (ms_dev * 1e-3), "ms_per_step": ms_dev / steps, "steps": steps}
            if e2e:
                for i in range(warmup):
                    self.step_e2e(i)
                _, ms_wall = self.timed(self.step_e2e, steps)
                out["e2e"] = world * steps / (ms_wall * 1e-3)
            return out

        def parity_outputs(self):
            """Frame 0 through the same network, eagerly, with the intermediate tensors fetched (for the parity leg)."""
            from oracle import parity
            fvr = FvRasterizer(self.net.fv_geometry) if self.views == 3 else None
            return parity.gpu_frame_outputs(self.net, BevRasterizer(**BEV), dev_frames[0][0], frames[0][1], im_info, calib, fvr)

        def close(self):
            self.pipe.drain(host_sync=True)
            del self.pipe, self.runner, self.net
            torch.cuda.empty_cache()

    leg = Leg(args.views, args.mode)
    with ClockSampler(local_rank) as clk:
        res = leg.measure(args.steps)
    value, e2e = res["value"], res["e2e"]
    launches = leg.launches_per_frame * args.steps
    h2d = sum(int(t.numel() * t.element_size()) for t in pin_frames[0])
    d2h = sum(int(t.numel() * t.element_size()) for t in leg.runner.host)

    # ---- roofline of the dominant kernel (conv/fc tcgen05 GEMM): per-launch CUDA events on the launch stream
    net, runner = leg.net, leg.runner
    kernels.GEMM_EVENTS = []
    net.use_side_stream = False   # serialise the two trunks so that every event pair brackets exactly one kernel
    for i in range(3):
        runner.load_device(*dev_frames[i % n_frames])
        runner._forward()             # eager replay of the same program (events cannot be read inside a graph)
    torch.cuda.synchronize()
    net.use_side_stream = True
    evs = kernels.GEMM_EVENTS
    kernels.GEMM_EVENTS = None
    gemm_ms = sum(ev[0].elapsed_time(ev[1]) for ev in evs) / 3.0
    n_gemm = len(evs) // 3
    per_kernel = {}
    for ev in evs:
        k = per_kernel.setdefault(ev[2], [0, 0.0, 0.0])
        k[0] += 1; k[1] += ev[0].elapsed_time(ev[1]); k[2] += ev[3]
    R = 300
    flops = gemm_flops(net, 701, 801, IMG_HW[0], IMG_HW[1], R)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_burst = float(peaks.get("bf16_tflops", 1645.0))
    all_tf = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    # the dominant kernel = the template instantiation with the largest share of the GEMM time
    dom = max(per_kernel.items(), key=lambda kv: kv[1][1])
    dom_name, (dom_n, dom_ms, dom_flops) = dom[0], dom[1]
    dom_tf = dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    traffic = None
    try:   # per-launch DRAM bytes of that kernel from the committed ncu capture of this same command
        tj = json.load(open(os.path.join(ROOT, "profiles", "gemm_dram_traffic.json")))
        rows = [r for r in tj["views%d" % args.views]["per_launch"] if dom_name.split("<")[0] in r["kernel"]
                and ("<%s," % dom_name.split("<")[1].split(",")[0]) in r["kernel"].replace(" ", "")]
        if rows:
            traffic = sum(r["dram_MB"] for r in rows) * 1e6 / len(rows)
    except Exception:
        pass
    mult = {"precise": 3, "mixed": 2, "fast": 1}[args.mode] if "<" not in dom_name else \
        {"3": 3, "2": 2, "1": 1}.get(dom_name.rstrip(">").split(",")[-1], 1)
    roofline = {"bound": "tensor", "achieved": dom_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": dom_tf / peak_tf,
                "frac_of_burst_peak": dom_tf / peak_burst, "traffic": traffic,
                "kernel": "%s: %d launches/frame, %.3f ms/frame, %.1f%% of the GEMM time; algorithmic FLOPs counted 1x "
                          "(the %s mode issues %dx the MMAs: %.0f TFLOP/s issued)"
                          % (dom_name, dom_n // 3, dom_ms / 3.0, 100.0 * dom_ms / (gemm_ms * 3.0), args.mode, mult, dom_tf * mult),
                "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (the kernels are timed inside a long step); "
                                "frac_of_burst_peak uses bf16_tflops") if peaks else "fallback 1.4 PFLOP/s",
                "all_gemm_kernels": {"launches_per_frame": n_gemm, "ms_per_frame": gemm_ms, "algorithmic_gflop_per_frame": flops / 1e9,
                                     "achieved": all_tf, "frac": all_tf / peak_tf,
                                     "by_kernel": {k: {"launches": v[0] // 3, "ms": v[1] / 3.0,
                                                       "tflops_1x": v[2] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0}
                                                   for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][1])}},
                "gemm_share_of_step": gemm_ms / res["ms_per_step"] if res["ms_per_step"] > 0 else None}

    line = {"metric": "MV3D inference frames/sec", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"precise": "bf16x3 (bf16 hi/lo split, 3 tcgen05 passes, fp32 accumulate)",
                      "mixed": "f16+2xe5m2 in the 3x3 convs and fc6 (fp16 pass + one e5m2 pass carrying both correction terms = 2 "
                               "pass-equivalents, fp32 accumulate), bf16x3 in the first image layer, 1x1 convs and the other fc layers",
                      "fast": "bf16"}[args.mode],
            "data": "synthetic", "config": make_config(args.views),
            "impl_config": {"mode": args.mode, "frames_in_flight": depth,
                            "execution": "one CUDA graph per frame (FrameRunner: the trunks on separate captured streams), "
                                         "independent batch-1 frames replayed round-robin on separate streams (FramePipeline)"},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clk.summary(), "roofline": roofline}

    # ---- parity, measured in THIS run: frame 0 of this network (and of the bf16x3 network below) against the CPU oracle
    # on the same weights.  GPU side here; the oracle runs in the cpu_baseline leg at the end.
    parity_gpu, cpu_params = {}, None
    if want_cpu:
        try:
            parity_gpu[args.mode] = leg.parity_outputs()
            cpu_params = {k: {kk: vv.cpu().numpy() for kk, vv in v.items()} for k, v in net.params.items()}
        except Exception as e:
            line["parity"] = {"error": repr(e)[:300]}
    leg.close()
    del net, runner

    # configs[2] right after the headline leg (its memory released): the second number BASELINE's metric names, measured
    # before the side legs below heat the board further
    if not args.no_train:
        try:
            line["train_step"] = run_train_bench(args, rank, world, local_rank)
        except Exception as e:  # the inference headline must survive a failure of the secondary leg
            line["train_step"] = {"error": repr(e)[:300]}
    if not args.no_other_views:
        # the other view count through the same pipeline, beside the headline (north_star names three views; the
        # reference has two).  Same clocks caveat: shorter run.
        try:
            leg = Leg(other_views, args.mode)
            n_alt = min(args.steps, 100)
            r2 = leg.measure(n_alt)
            line["views%d" % other_views] = {"value": r2["value"], "unit": "frames/s", "steps": n_alt,
                                             "e2e": {"value": r2["e2e"], "unit": "frames/s"}, "ms_per_step": r2["ms_per_step"],
                                             "config": make_config(other_views)}
            leg.close()
        except Exception as e:
            line["views%d" % other_views] = {"error": repr(e)[:300]}
    if args.mode == "mixed" and not args.no_precise_leg:
        # The same frames through the bf16 hi/lo 3-pass network (every GEMM at ~2^-17 per product), reported beside the
        # headline so that the cost of the tighter arithmetic is on record in the same run.
        try:
            leg = Leg(args.views, "precise")
            n_alt = min(args.steps, 100)
            r3 = leg.measure(n_alt, e2e=False)
            line["precise_mode"] = {"value": r3["value"], "unit": "frames/s", "steps": n_alt,
                                    "dtype": "bf16x3 (bf16 hi/lo split, 3 tcgen05 passes, fp32 accumulate) in every GEMM",
                                    "note": "device-resident inputs, same pipeline"}
            if want_cpu:
                parity_gpu["precise"] = leg.parity_outputs()
            leg.close()
        except Exception as e:
            line["precise_mode"] = {"error": repr(e)[:300]}
    if not args.no_extras:
        # configs[4] (raster sweep, every rank rasterises its own clouds: N GPUs = N x the frames) and configs[0]
        # (BEV-only 3D-RPN forward: the CPU port beside the device layer) in the driver-run line
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import raster_sweep
            rows, shape, peak_hbm = raster_sweep.sweep(iters=20, check=(world == 1), rank=rank, world=world)
            line["raster_sweep"] = {"workload": "configs[4]: LiDAR -> BEV 701x801x36 float32 raster, CUDA-graph replay of 20 frames, "
                                                "max over ranks; GB/s = (16 B/point + 4 B/output element) / time",
                                    "n_gpus": world, "hbm_peak_gbs": peak_hbm, "checked_bit_exact_vs_oracle": world == 1,
                                    "rows": rows}
        except Exception as e:
            line["raster_sweep"] = {"error": repr(e)[:300]}
        if want_cpu:
            try:
                import rpn_config0
                rows = rpn_config0.run(reps=1, gpu_iters=10)
                line["config0_rpn"] = {"workload": "configs[0]: BEV-only 3D-RPN forward (proposal_layer_3d: anchors + "
                                                   "bbox_transform + projection + filters + sort + NMS), CPU port on 1 host "
                                                   "thread (numpy + C NMS, per-box projection loop kept) vs the device layer",
                                       "rows": [dict(zip(("grid", "cfg", "pre_nms", "post_nms", "proposals", "cpu_ms", "b200_ms",
                                                          "survivors_identical"), r)) for r in rows]}
            except Exception as e:
                line["config0_rpn"] = {"error": repr(e)[:300]}
    if want_cpu:
        # the reference's CPU path on this box's host cores -- WITH THE GPU NETWORK'S WEIGHTS, so that the same oracle
        # frames also yield the parity numbers of this run
        params = cpu_params if cpu_params is not None else make_cpu_params(views=args.views)
        fps, ms, cores, _ = run_cpu_reference(2, 1, frames, args.views, params=params)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "2 whole frames after 1 warm-up (oracle port; conv/fc torch-CPU fp32)"}
        if "train_step" in line and "ms" in line.get("train_step", {}):
            try:
                one = cpu_train_step_ms(frames, params)
                line["train_step"]["cpu_baseline"] = {
                    "value": 2.0 * one, "unit": "ms per 2-frame step", "cores": cores, "kind": "port",
                    "sample": "ONE frame fwd+bwd timed (%.0f ms) x 2: the reference trains strictly batch 1, a configs[2] "
                              "step is two sequential CPU steps (torch-CPU fp32 autograd + numpy/C target layers; Adam "
                              "not included)" % one}
            except Exception as e:
                line["train_step"]["cpu_baseline"] = {"error": repr(e)[:200]}
        if parity_gpu and cpu_params is not None:
            try:
                from oracle import parity
                bv0 = orc.point_cloud_2_top(frames[0][0], **BEV)[None]
                fv0 = orc.point_cloud_2_front(frames[0][0])[None] if args.views == 3 else None
                ocfg = {"TEST": dict(RPN_PRE_NMS_TOP_N=6000, RPN_POST_NMS_TOP_N=300, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=5)}
                par = {"metric": "max|gpu - oracle| / max|oracle| per tensor (``_abs``: absolute); frame 0, %d views, same "
                                 "weights; ``_tf``: teacher-forced on the GPU's own conv5 maps and rois" % args.views,
                       "tolerance": parity.FLOAT_TOL}
                for mode, got in parity_gpu.items():
                    errs, exact, prop = parity.oracle_frame_errors(got, cpu_params, bv0, frames[0][1], im_info, calib,
                                                                   orc.CFG_GEOMETRY, cfg=ocfg, fv=fv0)
                    par[mode] = {"errors": errs, "exact": exact, "proposals": prop,
                                 "within_tolerance": bool(all(v < parity.FLOAT_TOL for v in errs.values()) and all(exact.values()))}
                line["parity"] = par
            except Exception as e:
                line["parity"] = {"error": repr(e)[:300]}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
