"""Worker of tests/test_gpu_multi.py (launched under torchrun, one process per GPU, NCCL):
  (1) inference: frames sharded over the ranks (sharding.assign), each rank runs its frames, rank 0 gathers the
      detections and compares them BITWISE with its own single-GPU run of all frames (SURVEY 8e / App. C last row);
  (2) training: each rank takes one train step on its own frames with the bucketed NCCL exchange; the all-reduced flat
      gradient must equal the sum of the per-rank single-GPU gradients (gathered raw) to 1e-6 of its max.
Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import build as ob
    ob.build()
    from oracle import mv3d_oracle as orc   # synthetic inputs only
    from mv3d_tf_b200 import sharding
    from mv3d_tf_b200.fast_rcnn.config import cfg, cfg_from_end2end_yml
    from mv3d_tf_b200.fast_rcnn.train_mv import SolverWrapper
    from mv3d_tf_b200.networks.factory import get_network
    from mv3d_tf_b200.utils.read_lidar import BevRasterizer
    from mv3d_tf_b200.utils.transform import BevGeometry

    cfg_from_end2end_yml()
    cfg.USE_GPU_NMS = False
    rk = dict(res=0.1, zres=0.3, side_range=(-8., 8.), fwd_range=(0., 16.), height_range=(-2, 0.4))
    geom_kw = dict(x_min=0, x_max=16, y_min=-8, y_max=8, res=0.1)
    im_info = np.array([[161, 161, 1]], np.float32)
    calib = np.array(orc.KITTI_CALIB, np.float32).copy()
    p2 = calib[0].reshape(3, 4); p2[0] *= 0.2; p2[1] *= 0.17; calib[0] = p2.reshape(-1)
    out = {}

    # ---------------- (1) inference determinism across ranks
    net = get_network("MV3D_test", bv_channels=9, precise=True, mixed=True, geometry=BevGeometry(**geom_kw), img_size=(64, 256))
    net.init_weights(seed=7, mode="he")
    raster = BevRasterizer(**rk)
    n_frames = 2 * world + 1

    def frame(i):
        pts = orc.synth_points(30000, seed=50 + i)
        pts[:, 0] *= 0.2
        pts[:, 1] *= 0.17
        img = np.random.default_rng(70 + i).normal(0, 50, (1, 64, 256, 3)).astype(np.float32)
        return pts, img

    def detect(i):
        pts, img = frame(i)
        bv = raster.to_pad(torch.from_numpy(pts).cuda(), precise=True)
        cls, box, rois = net.run([net.get_output("cls_prob"), net.get_output("bbox_pred"), net.get_output("roi_data_bv")],
                                 {net.lidar_bv_data: bv, net.image_data: img, net.im_info: im_info, net.calib: calib})
        n = int(net.last_num_rois.item())
        # fc6 is split-K with fp32 atomics (run-to-run order): the proposals are the bitwise-comparable part
        return torch.cat((torch.full((n, 1), float(i), device="cuda"), rois[:n, 1:]), 1)
    mine = sharding.assign(list(range(n_frames)), rank, world)
    rows = torch.cat([detect(i) for i in mine]) if mine else torch.zeros((0, 5), device="cuda")
    gathered = sharding.gather_rows(rows)
    if rank == 0:
        want = torch.cat([detect(i) for i in range(n_frames)])
        got = torch.cat(gathered)
        out["inference_rows"] = int(want.shape[0])
        out["inference_bitwise_equal"] = bool(got.shape == want.shape and torch.equal(got, want))
    del net
    torch.cuda.empty_cache()

    # ---------------- (2) all-reduced gradient == sum of the single-GPU gradients
    tnet = get_network("MV3D_train", bv_channels=9, precise=True, geometry=BevGeometry(**geom_kw), img_size=(64, 256))
    tnet.init_weights(seed=7, mode="he")
    ogeom = orc.BevGeometry(**geom_kw)

    def blobs(i):
        pts, img = frame(100 + i)
        gt = orc.synth_gt(4, seed=20 + i, geom=ogeom)
        return dict(lidar_bv_data=orc.point_cloud_2_top(pts, **rk)[None], image_data=img, im_info=im_info,
                    gt_boxes_bv=[gt[0]], gt_boxes_3d=[gt[1]], gt_boxes_corners=[gt[2]], calib=calib)
    sw = SolverWrapper(network=tnet, keep_prob=1.0, lr=1e-3, process_group=dist.group.WORLD)
    sw.exchange.bucket_bytes = 8 << 20           # several buckets even on this small problem
    ex = sw.exchange
    # record exactly what each bucket put on the wire (the backward-filter GEMMs combine their row splits with fp32
    # atomics, so two runs of the same step differ at the 1e-6 level: the comparison must use THIS run's local gradient)
    local_grad = torch.zeros_like(sw.grad)
    send = ex._send

    def recording_send(buf, lo, hi):
        if hi > lo:
            local_grad[lo:hi].copy_(buf[lo:hi])
        send(buf, lo, hi)
    ex._send = recording_send
    np.random.seed(3)
    sw.train_step(blobs(rank), keep_prob=1.0, apply_update=False)
    reduced = sw.grad.clone()
    parts = [torch.zeros_like(local_grad) for _ in range(world)]
    dist.all_gather(parts, local_grad)
    want = torch.stack(parts).double().sum(0)
    err = float((reduced.double() - want).abs().max() / want.abs().max())
    covered = bool((local_grad != 0).any()) and ex.buckets_last_step >= 2
    if rank == 0:
        out["grad_rel_err"] = err
        out["grad_buckets"] = ex.buckets_last_step
        out["grad_covered"] = covered
        out["grad_elems"] = int(reduced.numel())
        print("DISTCHECK " + json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
