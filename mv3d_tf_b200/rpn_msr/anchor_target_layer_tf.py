"""anchor_target_layer with the reference's py_func signature (lib/rpn_msr/anchor_target_layer_tf.py:21-250).

The O(N x G) float64 IoU, the two arg-max passes, the pre-sampling labels and the bbox_transform_3d targets run in
csrc/targets.cu.  The random sub-sampling (`npr.choice`, :146-159,176-183) is drawn from numpy's global RandomState
here on the host -- it must consume the same MT19937 stream as the reference to be reproducible -- on the compact
per-anchor code the kernel returns (one 35 kB device->host read per frame at the BASELINE shape).

  anchor_target_layer(...)      numpy in / numpy out, drop-in for the py_func
  AnchorTargetLayer(...)(...)   device tensors in / device tensors out
"""
from __future__ import annotations

import numpy as np
import numpy.random as npr
import torch

from .._lib import check, current_stream, lib, ptr
from ..fast_rcnn.config import cfg
from ..utils.transform import REF_GEOMETRY, BevGeometry, bv_anchor_to_lidar
from .generate_anchors import all_anchors


class AnchorTargetLayer:
    """One instance per feature-map size: caches the anchor tables (int32 boxes, float64 LiDAR boxes) on the device."""

    def __init__(self, height, width, feat_stride=8, geom: BevGeometry = REF_GEOMETRY, device="cuda"):
        self.Hf, self.Wf = int(height), int(width)
        self.device = torch.device(device)
        a = all_anchors(self.Hf, self.Wf, feat_stride)
        self.N = a.shape[0]
        self.anchors_h = a
        self.anchors3d_h = bv_anchor_to_lidar(a, geom)                      # float64 (N,6)   (:164)
        self.anchors = torch.from_numpy(a.astype(np.int32)).to(self.device)
        self.anchors3d = torch.from_numpy(np.ascontiguousarray(self.anchors3d_h, dtype=np.float64)).to(self.device)
        self.max_ov = torch.empty(self.N, dtype=torch.float64, device=self.device)
        self.argmax = torch.empty(self.N, dtype=torch.int32, device=self.device)
        self.code = torch.empty(self.N, dtype=torch.int8, device=self.device)
        self.code_h = torch.empty(self.N, dtype=torch.int8).pin_memory()
        self.labels_h = torch.empty(self.N, dtype=torch.float32).pin_memory()
        self.counts_h = torch.empty(2, dtype=torch.int32).pin_memory()

    def __call__(self, gt_boxes_bv: torch.Tensor, gt_boxes_3d: torch.Tensor, im_info, rng=None, want_rois=False):
        """gt_boxes_bv (G,5) / gt_boxes_3d (G,7) float32 CUDA.  Returns dict(labels (N,) f32, targets (N,6) f32,
        counts (2,) int32 = [#label != -1, #label == 1]) on the device (+ the <=128 sampled anchors when asked)."""
        c = cfg.TRAIN
        npr_ = npr if rng is None else rng
        G = gt_boxes_bv.shape[0]
        info = np.asarray(im_info, dtype=np.float32).reshape(-1, 3)[0]
        gt_bv = gt_boxes_bv.to(torch.float32).contiguous()
        gt_3d = gt_boxes_3d.to(torch.float32).contiguous()
        ws = torch.empty(G, dtype=torch.int64, device=self.device)
        targets = torch.empty((self.N, 6), dtype=torch.float32, device=self.device)
        check(lib().mv3d_anchor_targets(ptr(self.anchors), ptr(self.anchors3d), self.N, ptr(gt_bv), ptr(gt_3d), G,
                                        float(info[0]), float(info[1]), float(c.RPN_POSITIVE_OVERLAP),
                                        float(c.RPN_NEGATIVE_OVERLAP), int(bool(c.RPN_CLOBBER_POSITIVES)),
                                        ptr(self.max_ov), ptr(self.argmax), ptr(ws), ptr(self.code), ptr(targets),
                                        current_stream()), "mv3d_anchor_targets")
        self.code_h.copy_(self.code, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        code = self.code_h.numpy()
        inds_inside = np.where(code & 4)[0]
        labels = ((code[inds_inside] & 3).astype(np.float32) - 1.0)
        lt_neg = (code[inds_inside] & 8) != 0
        # ---- sub-sampling, exactly the reference's three draws (:146-159, :176-183)
        num_fg = int(c.RPN_FG_FRACTION * c.RPN_BATCHSIZE)
        fg_inds = np.where(labels == 1)[0]
        if len(fg_inds) > num_fg:
            labels[npr_.choice(fg_inds, size=(len(fg_inds) - num_fg), replace=False)] = -1
        num_bg = c.RPN_BATCHSIZE - np.sum(labels == 1)
        bg_inds = np.where(labels == 0)[0]
        if len(bg_inds) > num_bg:
            labels[npr_.choice(bg_inds, size=(len(bg_inds) - num_bg), replace=False)] = -1
        sampled = np.where(labels != -1)[0] if want_rois else None
        labels[lt_neg] = 0
        num_bg = c.RPN_BATCHSIZE - np.sum(labels == 1)
        bg_inds = np.where(labels == 0)[0]
        if len(bg_inds) > num_bg:
            labels[npr_.choice(bg_inds, size=(len(bg_inds) - num_bg), replace=False)] = -1
        full = self.labels_h.numpy()
        full.fill(-1)
        full[inds_inside] = labels
        self.counts_h[0] = int(np.sum(labels != -1))
        self.counts_h[1] = int(np.sum(labels == 1))
        out = dict(labels=self.labels_h.to(self.device, non_blocking=True),
                   counts=self.counts_h.to(self.device, non_blocking=True), targets=targets)
        if want_rois:
            idx = inds_inside[sampled]
            z = np.zeros((idx.shape[0], 1), np.float32)
            out["anchors"] = np.hstack((z, self.anchors_h[idx])).astype(np.float32)
            out["anchors_3d"] = np.hstack((z, self.anchors3d_h[idx])).astype(np.float32)
        return out


_layers = {}


def anchor_target_layer(rpn_cls_score, gt_boxes, gt_boxes_3d, im_info, _feat_stride=[8, ], anchor_scales=[1.0, 1.0]):
    """Drop-in for the py_func: (rpn_labels (N,), rpn_bbox_targets (N,6), anchors (<=128,5), anchors_3d (<=128,7))."""
    assert rpn_cls_score.shape[0] == 1, 'Only single item batches are supported'
    h, w = rpn_cls_score.shape[1:3]
    key = (h, w, int(_feat_stride[0]))
    layer = _layers.get(key)
    if layer is None:
        layer = _layers[key] = AnchorTargetLayer(h, w, int(_feat_stride[0]))
    gt_bv = torch.from_numpy(np.ascontiguousarray(gt_boxes, dtype=np.float32)).cuda()
    gt_3d = torch.from_numpy(np.ascontiguousarray(gt_boxes_3d, dtype=np.float32)).cuda()
    o = layer(gt_bv, gt_3d, im_info, want_rois=True)
    return o["labels"].cpu().numpy(), o["targets"].cpu().numpy(), o["anchors"], o["anchors_3d"]
