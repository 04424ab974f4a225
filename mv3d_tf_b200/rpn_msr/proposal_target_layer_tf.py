"""proposal_target_layer_3d with the reference's py_func signature (lib/rpn_msr/proposal_target_layer_tf.py:19-94,
_sample_rois_3d :227-298).  IoU / assignment and the output stage (corners, image projection, corner-regression
targets) run in csrc/targets.cu; the two `npr.choice` draws (:252,:267) stay on the host for the same reason as in
anchor_target_layer_tf.py.

  proposal_target_layer_3d(...)       numpy in / numpy out, drop-in for the py_func
  ProposalTargetLayer3D()(...)        device tensors in / device tensors out
"""
from __future__ import annotations

import numpy as np
import numpy.random as npr
import torch

from .._lib import check, current_stream, lib, ptr
from ..fast_rcnn.config import cfg
from ..utils.transform import projection_matrix


class ProposalTargetLayer3D:
    def __init__(self, device="cuda"):
        self.device = torch.device(device)

    def __call__(self, rois_bv: torch.Tensor, rois_3d: torch.Tensor, n_rois, gt_boxes_bv: torch.Tensor,
                 gt_boxes_3d: torch.Tensor, gt_boxes_corners: torch.Tensor, calib, num_classes: int,
                 batch_index: float = 0.0, rng=None):
        """rois_bv (cap,5) / rois_3d (cap,7) float32 CUDA with `n_rois` valid rows (int or int32 device tensor).
        Returns dict(bv (K,5), img (K,5), labels (K,) int32, targets (K,24*nc), p3d (K,7), n_fg)."""
        c = cfg.TRAIN
        npr_ = npr if rng is None else rng
        dev = self.device
        R = int(n_rois.item()) if isinstance(n_rois, torch.Tensor) else int(n_rois)
        G = gt_boxes_bv.shape[0]
        gt_bv = gt_boxes_bv.to(torch.float32).contiguous()
        gt_3d = gt_boxes_3d.to(torch.float32).contiguous()
        gt_cnr = gt_boxes_corners.to(torch.float32).contiguous()
        rois_bv, rois_3d = rois_bv.contiguous(), rois_3d.contiguous()
        max_ov = torch.empty(R + G, dtype=torch.float64, device=dev)
        assign = torch.empty(R + G, dtype=torch.int32, device=dev)
        check(lib().mv3d_roi_overlaps(ptr(rois_bv), R, ptr(gt_bv), G, ptr(max_ov), ptr(assign), current_stream()),
              "mv3d_roi_overlaps")
        mo = max_ov.cpu().numpy()   # synchronises
        rois_per_image = int(c.BATCH_SIZE) // 1                                     # :56
        fg_rois_per_image = np.round(c.FG_FRACTION * rois_per_image)                # :57
        fg_inds = np.where(mo >= c.FG_THRESH)[0]                                    # :244
        fg_n = int(min(fg_rois_per_image, fg_inds.size))
        if fg_inds.size > 0:
            fg_inds = npr_.choice(fg_inds, size=fg_n, replace=False)
        bg_inds = np.where((mo < c.BG_THRESH_HI) & (mo >= c.BG_THRESH_LO))[0]       # :258-259
        bg_n = min(rois_per_image - fg_n, bg_inds.size)
        if bg_inds.size > 0:
            bg_inds = npr_.choice(bg_inds, size=bg_n, replace=False)
        keep_h = np.append(fg_inds, bg_inds).astype(np.int32)                       # :272
        K = int(keep_h.shape[0])
        assert K > 0, "no roi sampled"
        keep = torch.from_numpy(keep_h).to(dev)
        proj = torch.from_numpy(projection_matrix(calib)).to(dev)
        out = dict(bv=torch.empty((K, 5), dtype=torch.float32, device=dev),
                   img=torch.empty((K, 5), dtype=torch.float32, device=dev),
                   labels=torch.empty((K,), dtype=torch.int32, device=dev),
                   targets=torch.empty((K, 24 * num_classes), dtype=torch.float32, device=dev),
                   p3d=torch.empty((K, 7), dtype=torch.float32, device=dev), n_fg=fg_n)
        check(lib().mv3d_proposal_targets(ptr(rois_bv), ptr(rois_3d), R, ptr(gt_bv), ptr(gt_3d), ptr(gt_cnr), G,
                                          ptr(keep), K, fg_n, ptr(assign), ptr(proj), int(num_classes),
                                          float(batch_index), ptr(out["bv"]), ptr(out["img"]), ptr(out["labels"]),
                                          ptr(out["targets"]), ptr(out["p3d"]), current_stream()),
              "mv3d_proposal_targets")
        return out


_layer = None


def proposal_target_layer_3d(rpn_rois_bv, rpn_rois_3d, gt_boxes_bv, gt_boxes_3d, gt_boxes_corners, calib, _num_classes):
    """Drop-in for the py_func: (rois_bv (K,5), rois_img (K,5), labels (K,1) int32, bbox_targets (K,24*nc), rois_3d (K,7))."""
    global _layer
    if _layer is None:
        _layer = ProposalTargetLayer3D()
    assert np.all(np.asarray(rpn_rois_bv)[:, 0] == 0), 'Only single item batches are supported'

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    o = _layer(dev(rpn_rois_bv), dev(rpn_rois_3d), int(np.asarray(rpn_rois_bv).shape[0]), dev(gt_boxes_bv),
               dev(gt_boxes_3d), dev(gt_boxes_corners), np.asarray(calib), int(_num_classes))
    return (o["bv"].cpu().numpy(), o["img"].cpu().numpy(), o["labels"].cpu().numpy().reshape(-1, 1),
            o["targets"].cpu().numpy(), o["p3d"].cpu().numpy())
