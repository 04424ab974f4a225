"""TEST INFRASTRUCTURE ONLY -- torch-CPU fp32 restatement of the TensorFlow part of the MV3D path.

PARITY UNPINNED: the arithmetic lives in TensorFlow 1.0 (README.md:9; not vendored, not installable here,
no golden vectors in the reference), so this file restates the *published semantics* of the TF ops at the
reference's call sites:
  tf.nn.conv2d SAME stride 1 + bias_add + relu   lib/networks/network.py:108-132
  tf.nn.max_pool 2x2/2 VALID                     :181-188
  fc = relu_layer / xw_plus_b, 4-D inputs flattened in (C,H,W) order   :369-397 (:381)
  softmax over the last dim                      :399-405
and the wiring of lib/networks/MV3D_test.py:32-123.  Weights are the reference's layout: conv HWIO, fc (in,out).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import mv3d_oracle as orc

TRUNK = [("conv1_1", 64), ("conv1_2", 64), "pool", ("conv2_1", 128), ("conv2_2", 128), "pool", ("conv3_1", 256),
         ("conv3_2", 256), ("conv3_3", 256), "pool", ("conv4_1", 512), ("conv4_2", 512), ("conv4_3", 512),
         ("conv5_1", 512), ("conv5_2", 512), ("conv5_3", 512)]


def _t(a, dtype):
    return torch.as_tensor(np.asarray(a), dtype=dtype)


def conv(x_nhwc, w_hwio, b, relu=True, dtype=torch.float32):
    """x (B,H,W,C) -> (B,H,W,Cout); SAME padding for 3x3, none for 1x1."""
    x = _t(x_nhwc, dtype).permute(0, 3, 1, 2)
    w = _t(w_hwio, dtype).permute(3, 2, 0, 1)
    y = F.conv2d(x, w, _t(b, dtype), padding=w.shape[-1] // 2)
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous()


def max_pool(x_nhwc):
    return F.max_pool2d(x_nhwc.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).contiguous()


def trunk(x_nhwc, params, suffix="", dtype=torch.float32, keep=None):
    x = _t(x_nhwc, dtype)
    for item in TRUNK:
        if item == "pool":
            x = max_pool(x)
        else:
            name = item[0] + suffix
            x = conv(x, params[name]["weights"], params[name]["biases"], True, dtype)
            if keep is not None:
                keep[name] = x
    return x


def fc(x, w, b, relu, dtype=torch.float32):
    x = _t(x, dtype)
    if x.dim() == 4:  # network.py:381: NHWC -> NCHW -> flatten
        x = x.permute(0, 3, 1, 2).reshape(x.shape[0], -1)
    y = x @ _t(w, dtype) + _t(b, dtype)
    return torch.relu(y) if relu else y


def rpn_head(conv5_3, params, dtype=torch.float32):
    """MV3D_test.py:70-83 -> (rpn_cls_prob_reshape (B,H,W,8), rpn_bbox_pred (B,H,W,24))."""
    r = conv(conv5_3, params["rpn_conv/3x3"]["weights"], params["rpn_conv/3x3"]["biases"], True, dtype)
    score = conv(r, params["rpn_cls_score"]["weights"], params["rpn_cls_score"]["biases"], False, dtype)
    bbox = conv(r, params["rpn_bbox_pred"]["weights"], params["rpn_bbox_pred"]["biases"], False, dtype)
    B, H, W, C = score.shape
    prob = torch.softmax(score.reshape(B, H, W * (C // 2), 2), dim=-1).reshape(B, H, W, C)
    return prob, bbox


def fusion_head(pool_bv, pool_img, params, dtype=torch.float32):
    """MV3D_test.py:103-123 -> (cls_prob (R,2), bbox_pred (R,48)); pooled inputs (R,7,7,512) NHWC."""
    f1 = fc(fc(pool_bv, params["fc6_1"]["weights"], params["fc6_1"]["biases"], True, dtype),
            params["fc7_1"]["weights"], params["fc7_1"]["biases"], True, dtype)
    f2 = fc(fc(pool_img, params["fc6_2"]["weights"], params["fc6_2"]["biases"], True, dtype),
            params["fc7_2"]["weights"], params["fc7_2"]["biases"], True, dtype)
    cat = torch.cat([f1, f2], dim=1)
    cls = torch.softmax(fc(cat, params["cls_score"]["weights"], params["cls_score"]["biases"], False, dtype), dim=-1)
    bbox = fc(cat, params["bbox_pred"]["weights"], params["bbox_pred"]["biases"], False, dtype)
    return cls, bbox


def mv3d_test_forward(bv, image, im_info, calib, params, cfg=None, geom=orc.REF_GEOMETRY, dtype=torch.float32,
                      keep=None, teacher=None):
    """Whole MV3D_test graph on the CPU.  `teacher` (optional dict) overrides stage inputs so that later stages
    can be compared on identical inputs: keys 'conv5_3', 'conv5_3_2', 'rois'."""
    teacher = teacher or {}
    with torch.no_grad():
        c5 = _t(teacher["conv5_3"], dtype) if "conv5_3" in teacher else trunk(bv, params, "", dtype, keep)
        c5_2 = _t(teacher["conv5_3_2"], dtype) if "conv5_3_2" in teacher else trunk(image, params, "_2", dtype, keep)
        prob, bbox = rpn_head(c5, params, dtype)
        if "rois" in teacher:
            rois_bv, rois_img, rois_3d = teacher["rois"]
        else:
            rois_bv, rois_img, rois_3d = orc.proposal_layer_3d(
                prob.float().numpy(), bbox.float().numpy(), np.asarray(im_info, np.float32), np.asarray(calib), "TEST",
                cfg=cfg, geom=geom)
        p1, _ = orc.roi_pool_fwd(c5.float().numpy(), rois_bv)
        p2, _ = orc.roi_pool_fwd(c5_2.float().numpy(), rois_img)
        cls, bb = fusion_head(p1, p2, params, dtype)
    return dict(conv5_3=c5, conv5_3_2=c5_2, rpn_cls_prob_reshape=prob, rpn_bbox_pred=bbox, rois_bv=rois_bv,
                rois_img=rois_img, rois_3d=rois_3d, pool_5=p1, pool_5_2=p2, cls_prob=cls, bbox_pred=bb)
