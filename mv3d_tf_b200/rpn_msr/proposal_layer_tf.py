"""proposal_layer_3d with the reference's py_func signature (lib/rpn_msr/proposal_layer_tf.py:25-202),
running as seven kernel launches in csrc/proposal.cu + csrc/nms.cu.

  proposal_layer_3d(...)            numpy in / numpy out, exact drop-in for the py_func
  ProposalLayer3D(...)(prob, deltas, calib)   device tensors in / device tensors out, no host sync
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .._lib import ProposalParams, check, current_stream, lib, ptr
from ..fast_rcnn.config import cfg
from ..utils.transform import REF_GEOMETRY, BevGeometry, bv_anchor_to_lidar, projection_matrix
from .generate_anchors import all_anchors, generate_anchors_bv


class ProposalLayer3D:
    """One instance per (feature map size, cfg_key, geometry): caches the anchor table and workspace."""

    def __init__(self, height, width, cfg_key="TEST", feat_stride=8, im_info=(601, 601, 1),
                 geom: BevGeometry = REF_GEOMETRY, img_size=(375, 1242), nms_rule_ge=None, device="cuda",
                 pre_nms_top_n=None, post_nms_top_n=None, nms_thresh=None, min_size=None):
        c = cfg[cfg_key]
        self.A = generate_anchors_bv().shape[0]
        self.Hf, self.Wf = int(height), int(width)
        self.N = self.Hf * self.Wf * self.A
        self.device = torch.device(device)
        p = ProposalParams()
        p.Hf, p.Wf, p.A = self.Hf, self.Wf, self.A
        p.xn, p.yn, p.x_min, p.y_min, p.res = geom.xn, geom.yn, geom.x_min, geom.y_min, geom.res
        p.im_h, p.im_w, p.im_scale = float(im_info[0]), float(im_info[1]), float(im_info[2])
        p.img_h, p.img_w = float(img_size[0]), float(img_size[1])   # hard-coded 375x1242 in the reference (:147)
        p.min_size = float(c.RPN_MIN_SIZE if min_size is None else min_size)
        p.pre_nms_top_n = int(c.RPN_PRE_NMS_TOP_N if pre_nms_top_n is None else pre_nms_top_n)
        p.post_nms_top_n = int(c.RPN_POST_NMS_TOP_N if post_nms_top_n is None else post_nms_top_n)
        p.nms_thresh = float(c.RPN_NMS_THRESH if nms_thresh is None else nms_thresh)
        # nms_wrapper.py:18-21: USE_GPU_NMS picks nms_kernel.cu's `>` rule, else cpu_nms.pyx's `>=`
        p.nms_rule_ge = int((not cfg.USE_GPU_NMS) if nms_rule_ge is None else nms_rule_ge)
        p.batch_index = 0.0
        self.params = p
        cap = self.N if p.pre_nms_top_n <= 0 else min(self.N, p.pre_nms_top_n)
        self.capacity = cap if p.post_nms_top_n <= 0 else min(cap, p.post_nms_top_n)
        a3d = bv_anchor_to_lidar(all_anchors(self.Hf, self.Wf, feat_stride), geom).astype(np.float32)
        self.anchors3d = torch.from_numpy(np.ascontiguousarray(a3d)).to(self.device)
        # scratch is taken per call from torch's caching allocator (free inside a captured graph): several frames may
        # be in flight on different streams through the same layer object
        self._ws_bytes = int(lib().mv3d_proposal_workspace_bytes(C.byref(p)))

    def __call__(self, prob: torch.Tensor, deltas: torch.Tensor, calib, batch_index: float = 0.0):
        """prob (Hf,Wf,2A) / deltas (Hf,Wf,6A) float32 CUDA.  `calib`: the (4,12) host array, or a float32 CUDA tensor
        with the 12 projection floats (P2.R0).Tr already on the device (CUDA-graph replay).  Returns dict of device
        tensors (capacity rows, rows >= num are zero) and `num` (int32[1], device)."""
        assert prob.is_cuda and prob.dtype == torch.float32 and prob.numel() == self.N * 2
        assert deltas.is_cuda and deltas.dtype == torch.float32 and deltas.numel() == self.N * 6
        # dense maps, or column ranges of one fused (Hf*Wf, 8A) head output: cells a fixed number of floats apart
        prob, self.params.ld_prob = self._cells(prob, 2 * self.A)
        deltas, self.params.ld_deltas = self._cells(deltas, 6 * self.A)
        if isinstance(calib, torch.Tensor):
            assert calib.is_cuda and calib.dtype == torch.float32 and calib.numel() == 12
            proj, self.params.d_proj = None, calib.data_ptr()
        else:
            proj, self.params.d_proj = projection_matrix(calib), None
        R, dev = self.capacity, self.device
        out = dict(bv=torch.empty((R, 5), dtype=torch.float32, device=dev),
                   img=torch.empty((R, 5), dtype=torch.float32, device=dev),
                   p3d=torch.empty((R, 7), dtype=torch.float32, device=dev),
                   scores=torch.empty((R,), dtype=torch.float32, device=dev),
                   anchor=torch.empty((R,), dtype=torch.int32, device=dev),
                   num=torch.empty((1,), dtype=torch.int32, device=dev))
        self.params.batch_index = float(batch_index)
        ws = torch.empty(self._ws_bytes, dtype=torch.uint8, device=dev)
        check(lib().mv3d_proposal_layer_3d(ptr(prob), ptr(deltas), ptr(self.anchors3d), ptr(proj),
                                           C.byref(self.params), ptr(out["bv"]), ptr(out["img"]), ptr(out["p3d"]),
                                           ptr(out["scores"]), ptr(out["anchor"]), ptr(out["num"]), ptr(ws),
                                           ws.numel(), current_stream()), "mv3d_proposal_layer_3d")
        return out

    def _cells(self, t: torch.Tensor, width: int):
        """(tensor, pitch): `t` viewed as Hf*Wf cells of `width` contiguous floats, `pitch` floats apart (0 = dense)."""
        t = t.reshape(self.Hf, self.Wf, width) if t.is_contiguous() else t
        if t.dim() == 3 and t.stride(2) == 1 and t.stride(0) == self.Wf * t.stride(1) and t.stride(1) >= width:
            return t, (0 if t.stride(1) == width else int(t.stride(1)))
        return t.contiguous(), 0

    def decode(self, prob: torch.Tensor, deltas: torch.Tensor, calib: np.ndarray):
        """Stage outputs of the decode kernel for every anchor (parity tests)."""
        dev, N = self.device, self.N
        score = torch.empty(N, dtype=torch.float32, device=dev)
        p3d = torch.empty((N, 6), dtype=torch.float32, device=dev)
        pbv = torch.empty((N, 4), dtype=torch.float32, device=dev)
        pimg = torch.empty((N, 4), dtype=torch.int32, device=dev)
        keep = torch.empty(N, dtype=torch.uint8, device=dev)
        proj = projection_matrix(calib)
        self.params.ld_prob = self.params.ld_deltas = 0
        check(lib().mv3d_proposal_decode(ptr(prob.contiguous()), ptr(deltas.contiguous()), ptr(self.anchors3d),
                                         ptr(proj), C.byref(self.params), ptr(score), ptr(p3d), ptr(pbv), ptr(pimg),
                                         ptr(keep), current_stream()), "mv3d_proposal_decode")
        return dict(score=score, p3d=p3d, pbv=pbv, pimg=pimg, keep=keep)


_layers = {}


def proposal_layer_3d(rpn_cls_prob_reshape, rpn_bbox_pred, im_info, calib, cfg_key, _feat_stride=[8, ],
                      anchor_scales=[1.0, 1.0]):
    """Drop-in for the py_func: returns (blob_bv (R,5), blob_img (R,5), blob_3d (R,7)) float32 numpy."""
    assert rpn_cls_prob_reshape.shape[0] == 1, 'Only single item batches are supported'
    info = np.asarray(im_info, dtype=np.float32).reshape(-1, 3)[0]
    h, w = rpn_cls_prob_reshape.shape[1:3]
    c = cfg[cfg_key]
    key = (h, w, cfg_key, int(_feat_stride[0]), tuple(info.tolist()), c.RPN_PRE_NMS_TOP_N, c.RPN_POST_NMS_TOP_N,
           c.RPN_NMS_THRESH, c.RPN_MIN_SIZE, bool(cfg.USE_GPU_NMS))
    layer = _layers.get(key)
    if layer is None:
        layer = _layers[key] = ProposalLayer3D(h, w, cfg_key, int(_feat_stride[0]), info)
    prob = torch.from_numpy(np.ascontiguousarray(rpn_cls_prob_reshape[0], dtype=np.float32)).cuda()
    deltas = torch.from_numpy(np.ascontiguousarray(rpn_bbox_pred[0], dtype=np.float32)).cuda()
    out = layer(prob, deltas, np.asarray(calib))
    n = int(out["num"].item())
    return out["bv"][:n].cpu().numpy(), out["img"][:n].cpu().numpy(), out["p3d"][:n].cpu().numpy()
