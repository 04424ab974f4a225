"""roi_pool / roi_pool_grad with the TF op's argument order (lib/roi_pooling_layer/roi_pooling_op.cc:30-49,
roi_pooling_op.py:4-7, roi_pooling_op_grad.py:23-43) over torch CUDA tensors."""
from __future__ import annotations

import torch

from .._lib import check, current_stream, lib, ptr


def roi_pool(bottom_data: torch.Tensor, bottom_rois: torch.Tensor, pooled_height: int, pooled_width: int,
             spatial_scale: float):
    """bottom_data (B,H,W,C) float32 NHWC, bottom_rois (R,5) -> (top_data (R,PH,PW,C) f32, argmax int32)."""
    if bottom_data.dim() != 4:
        raise ValueError("data must be 4-dimensional")       # roi_pooling_op.cc:83-85
    if bottom_rois.dim() != 2:
        raise ValueError("rois must be 2-dimensional")       # roi_pooling_op.cc:87-89
    assert bottom_data.is_cuda and bottom_data.dtype == torch.float32
    data = bottom_data.contiguous()
    rois = bottom_rois.to(torch.float32).contiguous()
    B, H, W, Cc = data.shape
    R = rois.shape[0]
    top = torch.empty((R, pooled_height, pooled_width, Cc), dtype=torch.float32, device=data.device)
    arg = torch.empty((R, pooled_height, pooled_width, Cc), dtype=torch.int32, device=data.device)
    check(lib().mv3d_roi_pool_forward(ptr(data), spatial_scale, R, H, W, Cc, pooled_height, pooled_width, ptr(rois),
                                      ptr(top), ptr(arg), current_stream()), "mv3d_roi_pool_forward")
    return top, arg


def roi_pool_grad(bottom_data: torch.Tensor, bottom_rois: torch.Tensor, argmax: torch.Tensor, grad: torch.Tensor,
                  pooled_height: int, pooled_width: int, spatial_scale: float) -> torch.Tensor:
    B, H, W, Cc = bottom_data.shape
    rois = bottom_rois.to(torch.float32).contiguous()
    out = torch.empty_like(bottom_data, dtype=torch.float32)
    check(lib().mv3d_roi_pool_backward(ptr(grad.contiguous()), spatial_scale, B, rois.shape[0], H, W, Cc,
                                       pooled_height, pooled_width, ptr(rois), ptr(out), ptr(argmax.contiguous()),
                                       current_stream()), "mv3d_roi_pool_backward")
    return out


class RoiPoolFunction(torch.autograd.Function):
    """autograd glue == @ops.RegisterGradient("RoiPool") (roi_pooling_op_grad.py:23-43): no gradient to rois."""

    @staticmethod
    def forward(ctx, data, rois, pooled_height, pooled_width, spatial_scale):
        top, arg = roi_pool(data, rois, pooled_height, pooled_width, spatial_scale)
        ctx.save_for_backward(data, rois, arg)
        ctx.cfg = (pooled_height, pooled_width, spatial_scale)
        ctx.mark_non_differentiable(arg)
        return top, arg

    @staticmethod
    def backward(ctx, grad_top, _grad_arg):
        data, rois, arg = ctx.saved_tensors
        ph, pw, sc = ctx.cfg
        return roi_pool_grad(data, rois, arg, grad_top, ph, pw, sc), None, None, None, None
