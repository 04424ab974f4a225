"""Inference entry points with the reference's names (lib/fast_rcnn/test_mv.py): box_detect, plus FrameRunner, the
production form of the same per-frame call: the whole frame (raster -> trunks -> RPN -> proposals -> ROI pool -> head)
captured once into a CUDA graph and replayed per frame with host buffers in / detections out.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from .. import kernels as K

from ..utils.read_lidar import BevRasterizer
from ..utils.transform import projection_matrix
from .config import cfg

FAR = -1.0e9   # x coordinate of padding points: outside every BEV crop, ignored by the rasteriser (read_lidar.py:58-62)


class FrameRunner:
    """One MV3D frame as a replayable CUDA graph.

    runner = FrameRunner(net, raster, max_points, img_hw, im_info)
    out = runner(points_pinned (n,4) f32, image_pinned (1,H,W,3) f32, calib (4,12))   # dict of pinned host tensors
    The graph owns static device input buffers (point cloud padded to `max_points` with far-away points, the image,
    the 12 projection floats) and static outputs; per frame the host does three async H2D copies, one graph launch and
    the D2H copies of the detections -- no per-kernel Python work and no tensor-map encoding on the critical path.
    """

    def __init__(self, net, raster: BevRasterizer, max_points: int, img_hw: Sequence[int], im_info,
                 fetch=("cls_prob", "bbox_pred", "roi_data_bv", "roi_data_img"), use_graph=True, device="cuda"):
        self.net, self.raster = net, raster
        self.device = torch.device(device)
        self.fv_raster = None
        if getattr(net, 'with_fv', False):   # third view: rasterised from the same (padded) cloud
            from ..utils.read_lidar import FvRasterizer
            self.fv_raster = FvRasterizer(net.fv_geometry, device=device)
        self.max_points = int(max_points)
        self.im_info = np.asarray(im_info, np.float32).reshape(-1, 3)
        self.pts = torch.full((self.max_points, 4), FAR, dtype=torch.float32, device=self.device)
        self.img = torch.zeros((1, int(img_hw[0]), int(img_hw[1]), 3), dtype=torch.float32, device=self.device)
        self.proj = torch.zeros(12, dtype=torch.float32, device=self.device)
        self.proj_h = torch.zeros(12, dtype=torch.float32).pin_memory()
        self.fetch_names = list(fetch)
        self.fetch = [net.get_output(f) for f in self.fetch_names]
        self.last_n = self.max_points
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.outs = None
        self.use_graph = use_graph
        self.host = None

    # ------------------------------------------------------------------
    def _forward(self):
        # mixed mode: the raster is written in the f16e5 operand format when conv1_1 can consume it (>= 17 channels pad to 64)
        fmt = K.FMT_F16E5 if (getattr(self.net, 'mixed', False) and self.raster.g["C"] > 16) else K.FMT_BF16X2
        bv = self.raster.to_pad(self.pts, precise=self.net.precise, fmt=fmt)
        feed = {self.net.lidar_bv_data: bv, self.net.image_data: self.img, self.net.im_info: self.im_info,
                self.net.calib: self.proj}
        if self.fv_raster is not None:
            feed[self.net.lidar_fv_data] = self.fv_raster(self.pts)[None]   # dense (1,H,W,3): conv1_1_3 takes the im2col path
        outs = self.net.run(self.fetch, feed)
        return list(outs) + [self.net.last_num_rois]

    def capture(self):
        """Warm up (packs weights, builds the proposal layer, sets kernel attributes) and capture the graph."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self.outs = self._forward()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if self.use_graph:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=s):
                self.outs = self._forward()
            torch.cuda.synchronize()
        self.host = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in self.outs]
        return self

    def replay(self):
        """Run the frame whose inputs are already in the static device buffers; returns the static device outputs."""
        if self.graph is None and self.outs is None:
            self.capture()
        if self.graph is not None:
            self.graph.replay()
        else:
            self.outs = self._forward()
        return self.outs

    def load_device(self, pts: torch.Tensor, img: torch.Tensor, calib=None):
        """Inputs (device tensors, or pinned host tensors) -> static buffers, asynchronously on the current stream: the
        caller's buffers must stay untouched until the frame's results were collected / the stream synchronised."""
        n = pts.shape[0]
        assert n <= self.max_points
        self.pts[:n].copy_(pts[:, :4], non_blocking=True)
        if n < self.last_n or self.last_n < self.max_points and n != self.last_n:
            if n < self.max_points:
                self.pts[n:].fill_(FAR)
        self.last_n = n
        self.img.copy_(img.view(self.img.shape), non_blocking=True)
        if calib is not None:
            self.proj_h.copy_(torch.from_numpy(projection_matrix(calib).reshape(-1)))
            self.proj.copy_(self.proj_h, non_blocking=True)

    def __call__(self, points: torch.Tensor, image: torch.Tensor, calib):
        """Host (ideally pinned) buffers in -> dict of pinned host tensors out; synchronises the current stream once."""
        if self.host is None:
            self.capture()
        self.load_device(points, image, calib)
        outs = self.replay()
        for h, o in zip(self.host, outs):
            h.copy_(o, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        res = dict(zip(self.fetch_names + ["num_rois"], self.host))
        return res


class FramePipeline:
    """`depth` FrameRunners (each its own CUDA graph and static buffers, same network weights) replayed round-robin on
    `depth` streams: while one frame sits in its serial tail (proposal sort, the NMS keep-chain, weight-streaming fc6)
    the next frame's trunks fill the SMs.  Every frame is still an independent batch-1 inference.

        pipe = FramePipeline(lambda: FrameRunner(net, BevRasterizer(...), ...), depth=2)
        pipe.submit(points_pinned, image_pinned, calib)      # returns immediately
        out = pipe.collect()                                  # oldest frame in flight: dict of pinned host tensors
    """

    def __init__(self, make_runner, depth=2):
        self.runners = [make_runner().capture() for _ in range(depth)]
        self.streams = [torch.cuda.Stream() for _ in range(depth)]
        self.done = [None] * depth
        self.head = 0      # next slot to submit into
        self.tail = 0      # oldest slot not yet collected
        self.depth = depth

    def submit(self, points, image, calib=None, device_inputs=False):
        k = self.head % self.depth
        assert self.head - self.tail < self.depth, "collect() the oldest frame first"
        r, st = self.runners[k], self.streams[k]
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            r.load_device(points, image, calib)
            outs = r.replay()
            if not device_inputs:
                for h, o in zip(r.host, outs):
                    h.copy_(o, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(st)
        self.done[k] = ev
        self.head += 1

    def collect(self):
        assert self.tail < self.head, "nothing in flight"
        k = self.tail % self.depth
        self.done[k].synchronize()
        self.tail += 1
        r = self.runners[k]
        return dict(zip(r.fetch_names + ["num_rois"], r.host))

    def drain(self, host_sync=False):
        """Make the current stream wait for everything in flight (device-side join).  With `host_sync` (or whenever a
        frame's results were not collected) the per-slot events are also waited for on the host, because the next
        submit() rewrites that slot's pinned calib / output buffers while its async copies could still be running."""
        for st in self.streams:
            torch.cuda.current_stream().wait_stream(st)
        if host_sync or self.tail < self.head:
            for ev in self.done:
                if ev is not None:
                    ev.synchronize()
        self.tail = self.head


def box_detect(sess, net, im, bv, calib, boxes=None):
    """test_mv.py:149-264: one frame -> (scores (R,2), pred_boxes_bv (R,8), pred_boxes_cnr (R,48), pred_boxes_cnr_r
    (R,48)).  `im` is the raw BGR image (PIXEL_MEANS are subtracted here, :163), `bv` the (H,W,C) BEV map.
    As in the reference, the reported boxes are the un-regressed proposals duplicated per class (:253-255) and the
    corner regression is returned separately."""
    im_blob = (np.asarray(im, np.float32) - cfg.PIXEL_MEANS.astype(np.float32))[None]
    bv_blob = np.asarray(bv, np.float32)[None]
    im_info = np.array([[bv_blob.shape[1], bv_blob.shape[2], 1]], dtype=np.float32)
    feed = {net.lidar_bv_data: bv_blob, net.image_data: im_blob, net.im_info: im_info, net.calib: np.asarray(calib),
            net.keep_prob: 1.0}
    cls_prob, bbox_pred, rois = net.run([net.get_output('cls_prob'), net.get_output('bbox_pred'),
                                         net.get_output('rois')], feed)
    n = int(net.last_num_rois.item())
    scores = cls_prob[:n].cpu().numpy()
    deltas = bbox_pred[:n].cpu().numpy()
    boxes_3d = rois['p3d'][:n, 1:7].cpu().numpy()
    boxes_cnr = lidar_3d_to_corners(boxes_3d)
    pred_boxes_cnr = np.hstack((boxes_cnr, boxes_cnr))
    pred_boxes_cnr_r = bbox_transform_inv_cnr(boxes_cnr, deltas)
    pred_boxes_bv = corners_to_bv(pred_boxes_cnr, net.geometry)
    return scores, pred_boxes_bv, pred_boxes_cnr, pred_boxes_cnr_r


# ---- post-processing helpers of box_detect (host side, O(R) with R <= 300) ---------------------------
def lidar_3d_to_corners(p):
    """lib/utils/transform.py:290-315 -> (N,24) [x0..7, y0..7, z0..7] in the input dtype."""
    sx = np.array([1, 1, -1, -1, 1, 1, -1, -1]); sy = np.array([1, -1, -1, 1, 1, -1, -1, 1]); sz = np.array([-1] * 4 + [1] * 4)
    dt = p.dtype
    hl, hw, hh = p[:, 3:4] / 2., p[:, 4:5] / 2., p[:, 5:6] / 2.
    xs = np.where(sx > 0, hl, -hl).astype(dt) + p[:, 0:1]
    ys = np.where(sy > 0, hw, -hw).astype(dt) + p[:, 1:2]
    zs = np.where(sz > 0, hh, -hh).astype(dt) + p[:, 2:3]
    return np.hstack((xs, ys, zs)).astype(dt, copy=False)


def bbox_transform_inv_cnr(boxes, deltas):
    """lib/fast_rcnn/bbox_transform.py:157-176: corners + deltas * diagonal, per class block of 24."""
    if boxes.shape[0] == 0:
        return np.zeros((0, deltas.shape[1]), dtype=deltas.dtype)
    boxes = boxes.astype(deltas.dtype, copy=False)
    diag = np.linalg.norm(boxes[:, 0::8] - boxes[:, 6::8], axis=1).reshape(-1, 1)
    n_cls = deltas.shape[1] // 24
    pred = np.zeros(deltas.shape, dtype=deltas.dtype)
    for k in range(n_cls):
        pred[:, 24 * k:24 * k + 24] = deltas[:, 24 * k:24 * k + 24] * diag + boxes
    return pred


def corners_to_bv(corners, geom):
    """lib/utils/transform.py:342-366: per class block of 24 corner coords -> (x1,y1,x2,y2) BEV box."""
    n_cls = corners.shape[1] // 24
    out = np.zeros((corners.shape[0], 4 * n_cls))   # float64 like the reference's np.zeros (transform.py:361)
    for k in range(n_cls):
        c = corners[:, 24 * k:24 * k + 24]
        xs, ys = c[:, 0:8], c[:, 8:16]
        xmax, xmin = xs.max(axis=1), xs.min(axis=1)
        ymax, ymin = ys.max(axis=1), ys.min(axis=1)
        # transform.py:13-20 on the corners' own dtype (float32 arrays with Python-float constants stay float32)
        out[:, 4 * k + 0] = geom.yn - (ymax - geom.y_min) // geom.res
        out[:, 4 * k + 1] = geom.xn - (xmax - geom.x_min) // geom.res
        out[:, 4 * k + 2] = geom.yn - (ymin - geom.y_min) // geom.res
        out[:, 4 * k + 3] = geom.xn - (xmin - geom.x_min) // geom.res
    return out


def collect_detections(scores, boxes_bv, boxes_cnr, num_classes, thresh=0.05, nms_thresh=None, max_per_image=300):
    """The per-frame tail of test_net (test_mv.py:420-444, 492-501): per foreground class keep scores > thresh, NMS on
    the BEV boxes with cfg.TEST.NMS (lib/utils/nms.pyx:17-68, the `>=` rule -- run on the GPU by nms.cpu_nms), then cap
    the frame at `max_per_image` detections over all classes.  Returns ({cls: (n,5) [x1,y1,x2,y2,score]},
    {cls: (n,25) [24 corner coords, score]})."""
    from ..nms.gpu_nms import cpu_nms as nms   # utils.cython_nms.nms == cpu_nms.pyx arithmetic

    nms_thresh = cfg.TEST.NMS if nms_thresh is None else nms_thresh
    dets, dets_cnr = {}, {}
    for j in range(1, num_classes):
        inds = np.where(scores[:, j] > thresh)[0]
        cls_scores = scores[inds, j]
        cls_dets = np.hstack((boxes_bv[inds, j * 4:(j + 1) * 4], cls_scores[:, np.newaxis])).astype(np.float32, copy=False)
        cls_dets_cnr = np.hstack((boxes_cnr[inds, j * 24:(j + 1) * 24], cls_scores[:, np.newaxis])).astype(np.float32, copy=False)
        keep = nms(cls_dets, nms_thresh)
        dets[j] = cls_dets[keep, :]
        dets_cnr[j] = cls_dets_cnr[keep, :]
    if max_per_image > 0 and num_classes > 1:
        image_scores = np.hstack([dets[j][:, -1] for j in range(1, num_classes)])
        if len(image_scores) > max_per_image:
            image_thresh = np.sort(image_scores)[-max_per_image]
            for j in range(1, num_classes):
                keep = np.where(dets[j][:, -1] >= image_thresh)[0]
                dets[j] = dets[j][keep, :]
                dets_cnr[j] = dets_cnr[j][keep, :]
    return dets, dets_cnr


def test_net(sess, net, imdb, weights_filename=None, max_per_image=300, thresh=0.05, vis=False):
    """test_mv.py:321-517 without the disk / plotting parts: `imdb` needs `num_classes`, `image_index` and
    `frame_at(i) -> (image HxWx3 raw BGR, bv HxWxC, calib 4x12)` (the reference reads the three from disk,
    :398-403).  Returns (all_boxes[cls][image], all_boxes_cnr[cls][image]) as the reference pickles them (:503-509)."""
    n = len(imdb.image_index)
    all_boxes = [[[] for _ in range(n)] for _ in range(imdb.num_classes)]
    all_boxes_cnr = [[[] for _ in range(n)] for _ in range(imdb.num_classes)]
    for i in range(n):
        im, bv, calib = imdb.frame_at(i)
        scores, boxes_bv, boxes_cnr, _ = box_detect(sess, net, im, bv, calib)
        dets, dets_cnr = collect_detections(scores, boxes_bv, boxes_cnr, imdb.num_classes, thresh=thresh,
                                            max_per_image=max_per_image)
        for j in range(1, imdb.num_classes):
            all_boxes[j][i] = dets[j]
            all_boxes_cnr[j][i] = dets_cnr[j]
    return all_boxes, all_boxes_cnr
