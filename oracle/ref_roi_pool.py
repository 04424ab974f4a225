"""TEST INFRASTRUCTURE ONLY.  ctypes wrapper over oracle/_ref/libref_roi_pool.so = the reference's own
lib/roi_pooling_layer/roi_pooling_op.cc (RoiPoolOp / RoiPoolGradOp, CPU kernels) compiled unmodified against the
stand-in TensorFlow headers of oracle/tf_stub/ (recipe: oracle/build_ref_roi_pool.py)."""
import ctypes
import os

import numpy as np

from . import build_ref_roi_pool

_LIB = None


def available() -> bool:
    return build_ref_roi_pool.build() is not None


def _lib():
    global _LIB
    if _LIB is None:
        path = build_ref_roi_pool.build()
        if path is None:
            raise RuntimeError("reference RoiPool binary not built and /root/reference absent")
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def roi_pool_forward(data, rois, pooled_h=7, pooled_w=7, spatial_scale=0.125):
    """RoiPool (roi_pooling_op.cc:74-190): data (B,H,W,C) f32, rois (R,5) f32 -> (top (R,PH,PW,C) f32, argmax i32)."""
    data = np.ascontiguousarray(data, np.float32)
    rois = np.ascontiguousarray(rois, np.float32)
    B, H, W, C = data.shape
    R = rois.shape[0]
    top = np.empty((R, pooled_h, pooled_w, C), np.float32)
    arg = np.empty((R, pooled_h, pooled_w, C), np.int32)
    rc = _lib().ref_roi_pool_forward(_p(data), B, H, W, C, _p(rois), R, pooled_h, pooled_w, ctypes.c_float(spatial_scale),
                                     _p(top), _p(arg))
    assert rc == 0, rc
    return top, arg


def roi_pool_backward(data, rois, argmax, grad, pooled_h=7, pooled_w=7, spatial_scale=0.125):
    """RoiPoolGrad (roi_pooling_op.cc:319-452) -> d(data) (B,H,W,C) f32."""
    data = np.ascontiguousarray(data, np.float32)
    rois = np.ascontiguousarray(rois, np.float32)
    argmax = np.ascontiguousarray(argmax, np.int32)
    grad = np.ascontiguousarray(grad, np.float32)
    B, H, W, C = data.shape
    out = np.empty(data.shape, np.float32)
    rc = _lib().ref_roi_pool_backward(_p(data), B, H, W, C, _p(rois), rois.shape[0], pooled_h, pooled_w,
                                      ctypes.c_float(spatial_scale), _p(argmax), _p(grad), _p(out))
    assert rc == 0, rc
    return out
