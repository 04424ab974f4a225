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


def fusion_head(pool_bv, pool_img, params, dtype=torch.float32, pool_fv=None):
    """MV3D_test.py:103-123 -> (cls_prob (R,2), bbox_pred (R,48)); pooled inputs (R,7,7,512) NHWC.  `pool_fv` adds the
    front-view branch fc6_3/fc7_3 (this project's extension; the reference has two branches)."""
    f1 = fc(fc(pool_bv, params["fc6_1"]["weights"], params["fc6_1"]["biases"], True, dtype),
            params["fc7_1"]["weights"], params["fc7_1"]["biases"], True, dtype)
    f2 = fc(fc(pool_img, params["fc6_2"]["weights"], params["fc6_2"]["biases"], True, dtype),
            params["fc7_2"]["weights"], params["fc7_2"]["biases"], True, dtype)
    feats = [f1, f2]
    if pool_fv is not None:
        feats.append(fc(fc(pool_fv, params["fc6_3"]["weights"], params["fc6_3"]["biases"], True, dtype),
                        params["fc7_3"]["weights"], params["fc7_3"]["biases"], True, dtype))
    cat = torch.cat(feats, dim=1)
    cls = torch.softmax(fc(cat, params["cls_score"]["weights"], params["cls_score"]["biases"], False, dtype), dim=-1)
    bbox = fc(cat, params["bbox_pred"]["weights"], params["bbox_pred"]["biases"], False, dtype)
    return cls, bbox


def mv3d_test_forward(bv, image, im_info, calib, params, cfg=None, geom=orc.REF_GEOMETRY, dtype=torch.float32,
                      keep=None, teacher=None, fv=None, fv_geom=orc.FV_GEOMETRY):
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
        out = {}
        p3 = None
        if fv is not None:   # front-view branch (extension): third trunk, FV rois from the 3-D proposals
            c5_3 = _t(teacher["conv5_3_3"], dtype) if "conv5_3_3" in teacher else trunk(fv, params, "_3", dtype, keep)
            rois_fv = np.hstack((np.asarray(rois_3d)[:, :1], orc.lidar_3d_to_fv(np.asarray(rois_3d)[:, 1:7], fv_geom)))
            p3, _ = orc.roi_pool_fwd(c5_3.float().numpy(), rois_fv.astype(np.float32))
            out.update(conv5_3_3=c5_3, rois_fv=rois_fv.astype(np.float32), pool_5_3=p3)
        cls, bb = fusion_head(p1, p2, params, dtype, pool_fv=p3)
    out.update(conv5_3=c5, conv5_3_2=c5_2, rpn_cls_prob_reshape=prob, rpn_bbox_pred=bbox, rois_bv=rois_bv,
               rois_img=rois_img, rois_3d=rois_3d, pool_5=p1, pool_5_2=p2, cls_prob=cls, bbox_pred=bb)
    return out


# ----------------------------------------------------------------------------------------------------------------
# training step (PARITY UNPINNED for the TF part, see the header): MV3D_train.py:42-182 wiring, the four losses of
# train_mv.py:94-139, gradients by torch autograd, Adam in TensorFlow's formulation (train_mv.py:144-146).
# ----------------------------------------------------------------------------------------------------------------
class _RoiPool(torch.autograd.Function):
    """RoiPool / RoiPoolGrad (roi_pooling_op.cc:123-182,369-444) through the C restatement in oracle_c.c."""

    @staticmethod
    def forward(ctx, data, rois):
        top, arg = orc.roi_pool_fwd(data.detach().float().numpy(), np.asarray(rois, np.float32))
        ctx.shape = tuple(data.shape)
        ctx.rois = np.asarray(rois, np.float32)
        ctx.arg = arg
        ctx.dtype = data.dtype
        return torch.as_tensor(top).to(data.dtype)

    @staticmethod
    def backward(ctx, g):
        d = orc.roi_pool_bwd(ctx.shape, ctx.rois, ctx.arg, np.ascontiguousarray(g.detach().float().numpy()))
        return torch.as_tensor(d).to(ctx.dtype), None


class _GatedReLU(torch.autograd.Function):
    """ReLU whose on/off pattern is given (teacher-forced discrete decision): y = x * mask, dy/dx = mask."""

    @staticmethod
    def forward(ctx, x, mask):
        ctx.save_for_backward(mask)
        return x * mask

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return g * mask, None


class _PoolAt(torch.autograd.Function):
    """2x2/2 max-pool whose arg-max positions are given: idx (B,C,Ho,Wo) flat indices into H*W."""

    @staticmethod
    def forward(ctx, x, idx):
        ctx.save_for_backward(idx)
        ctx.shape = x.shape
        return x.flatten(2).gather(2, idx.flatten(2)).view(idx.shape)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        B, C, H, W = ctx.shape
        out = torch.zeros((B, C, H * W), dtype=g.dtype)
        out.scatter_add_(2, idx.flatten(2), g.flatten(2))
        return out.view(B, C, H, W), None


class _RoiPoolAt(torch.autograd.Function):
    """ROI max-pool whose arg-max (flat index (h*W+w)*C+c inside the roi's frame, -1 = empty bin) is given."""

    @staticmethod
    def forward(ctx, data, rois, arg):
        B, H, W, C = data.shape
        b = torch.as_tensor(np.asarray(rois)[:, 0]).long().view(-1, 1, 1, 1)
        a = torch.as_tensor(np.asarray(arg)).long()
        flat = data.reshape(B, -1)
        top = flat[b.expand_as(a), a.clamp_min(0)] * (a >= 0)
        ctx.shape, ctx.rois, ctx.arg, ctx.dtype = tuple(data.shape), np.asarray(rois, np.float32), np.asarray(arg), data.dtype
        return top

    @staticmethod
    def backward(ctx, g):
        d = orc.roi_pool_bwd(ctx.shape, ctx.rois, ctx.arg, np.ascontiguousarray(g.detach().float().numpy()))
        return torch.as_tensor(d).to(ctx.dtype), None, None


def smooth_l1(d, sigma=3.0):
    """train_mv.py:67-84."""
    s2 = sigma * sigma
    sign = (d.abs() < 1.0 / s2).to(d.dtype)
    return d * d * 0.5 * s2 * sign + (d.abs() - 0.5 / s2) * (sign - 1.0).abs()


def train_forward_backward(bv, image, im_info, calib, gt_bv, gt_3d, gt_cnr, params, geom=orc.REF_GEOMETRY, cfg=None,
                           dtype=torch.float32, teacher=None, train_cfg=None, gates=None):
    """One frame: forward of MV3D_train, the four losses, backward.  Returns (losses dict, grads dict in the reference
    variable layouts, stage dict).  `teacher` may pin 'rpn_data' = (labels, targets) and 'roi_data' =
    (rois_bv, rois_img, labels, targets); `gates` may pin the remaining discrete decisions of the graph -- 'relu':
    {layer: bool mask, NHWC / (R,n)}, 'pool': {'pool1'+suffix: (B,C,Ho,Wo) indices}, 'roi': {'pool_5': argmax,
    'pool_5_2': argmax} -- so that gradients are compared on identical discrete decisions (a ReLU whose
    pre-activation is within the forward tolerance of zero would otherwise flip and move the gradient by
    O(1/sqrt(#units)), far above any float tolerance)."""
    teacher = teacher or {}
    gates = gates or {}
    P = {k: {kk: _t(vv, dtype).clone().requires_grad_(True) for kk, vv in v.items()} for k, v in params.items()}

    def act(y, name, nchw):
        m = gates.get("relu", {}).get(name)
        if m is None:
            return torch.relu(y)
        m = torch.as_tensor(np.asarray(m)).to(dtype)
        return _GatedReLU.apply(y, m.permute(0, 3, 1, 2) if nchw else m)

    def cv(x, name, relu=True):
        w = P[name]["weights"].permute(3, 2, 0, 1)
        y = F.conv2d(x, w, P[name]["biases"], padding=w.shape[-1] // 2)
        return act(y, name, True) if relu else y

    def tr(x, suffix):
        x = _t(x, dtype).permute(0, 3, 1, 2)
        n_pool = 0
        for item in TRUNK:
            if item == "pool":
                n_pool += 1
                idx = gates.get("pool", {}).get("pool%d%s" % (n_pool, suffix))
                x = F.max_pool2d(x, 2, 2) if idx is None else _PoolAt.apply(x, torch.as_tensor(np.asarray(idx)).long())
            else:
                x = cv(x, item[0] + suffix)
        return x

    c5, c5_2 = tr(bv, ""), tr(image, "_2")
    r = cv(c5, "rpn_conv/3x3")
    score = cv(r, "rpn_cls_score", False).permute(0, 2, 3, 1)       # (1,H,W,8)
    bbox = cv(r, "rpn_bbox_pred", False).permute(0, 2, 3, 1)        # (1,H,W,24)
    B, H, W, C = score.shape
    if "rpn_data" in teacher:
        labels, targets = teacher["rpn_data"]
    else:
        labels, targets, _, _ = orc.anchor_target_layer(score.detach().float().numpy(), gt_bv, gt_3d, im_info, geom=geom,
                                                        cfg=train_cfg)
    if "roi_data" in teacher:
        rois_bv, rois_img, rlab, rtgt = teacher["roi_data"]
    else:
        prob = torch.softmax(score.detach().reshape(B, H, W * (C // 2), 2), dim=-1).reshape(B, H, W, C)
        with np.errstate(all="ignore"):
            rb, _, r3 = orc.proposal_layer_3d(prob.float().numpy(), bbox.detach().float().numpy(),
                                              np.asarray(im_info, np.float32), np.asarray(calib), "TRAIN", cfg=cfg, geom=geom)
        rois_bv, rois_img, rlab, rtgt, _ = orc.proposal_target_layer_3d(rb, r3, gt_bv, gt_3d, gt_cnr, calib, 2, cfg=train_cfg)
    labels_t = torch.as_tensor(np.asarray(labels).reshape(-1))
    keep = labels_t != -1
    pos = labels_t == 1
    sc2 = score.reshape(-1, 2)
    rpn_ce = F.cross_entropy(sc2[keep], labels_t[keep].long())                               # :97-105
    d = bbox.reshape(-1, 6)[pos] - _t(np.asarray(targets).reshape(-1, 6), dtype)[pos]
    rpn_box = smooth_l1(d).sum(dim=1).mean()                                                 # :108-119
    ra = gates.get("roi", {})
    d1, d2 = c5.permute(0, 2, 3, 1).contiguous(), c5_2.permute(0, 2, 3, 1).contiguous()
    p1 = _RoiPoolAt.apply(d1, rois_bv, ra["pool_5"]) if "pool_5" in ra else _RoiPool.apply(d1, rois_bv)
    p2 = _RoiPoolAt.apply(d2, rois_img, ra["pool_5_2"]) if "pool_5_2" in ra else _RoiPool.apply(d2, rois_img)

    def fcl(x, name, relu):
        if x.dim() == 4:
            x = x.permute(0, 3, 1, 2).reshape(x.shape[0], -1)
        y = x @ P[name]["weights"] + P[name]["biases"]
        return act(y, name, False) if relu else y
    f1 = fcl(fcl(p1, "fc6_1", True), "fc7_1", True)
    f2 = fcl(fcl(p2, "fc6_2", True), "fc7_2", True)
    cat = torch.cat([f1, f2], dim=1)                                                         # keep_prob = 1
    cls_score, bbox_pred = fcl(cat, "cls_score", False), fcl(cat, "bbox_pred", False)
    ce = F.cross_entropy(cls_score, torch.as_tensor(np.asarray(rlab).reshape(-1)).long())    # :123-125
    box = smooth_l1(bbox_pred - _t(rtgt, dtype)).sum(dim=1).mean()                           # :128-133
    loss = ce + box + rpn_ce + rpn_box                                                       # :136
    loss.backward()
    grads = {k: {kk: (vv.grad if vv.grad is not None else torch.zeros_like(vv)).detach().double().numpy()
                 for kk, vv in v.items()} for k, v in P.items()}
    losses = dict(rpn_loss_cls=float(rpn_ce), rpn_loss_box=float(rpn_box), loss_cls=float(ce), loss_box=float(box))
    stage = dict(rpn_labels=np.asarray(labels), rpn_targets=np.asarray(targets), rois_bv=rois_bv, rois_img=rois_img,
                 roi_labels=np.asarray(rlab), roi_targets=np.asarray(rtgt), cls_score=cls_score.detach(),
                 bbox_pred=bbox_pred.detach(), rpn_cls_score=score.detach(), rpn_bbox_pred=bbox.detach(),
                 conv5_3=c5.detach().permute(0, 2, 3, 1))
    return losses, grads, stage


def adam_step(params, grads, m, v, step, lr=1e-5, b1=0.9, b2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer (TF 1.0 defaults): epsilon outside the bias correction.  In place on float64 dicts."""
    lr_t = lr * np.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
    for k in params:
        for kk in params[k]:
            g = grads[k][kk]
            m[k][kk] = b1 * m[k][kk] + (1 - b1) * g
            v[k][kk] = b2 * v[k][kk] + (1 - b2) * g * g
            params[k][kk] = params[k][kk] - lr_t * m[k][kk] / (np.sqrt(v[k][kk]) + eps)
