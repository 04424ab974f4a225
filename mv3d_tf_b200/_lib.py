"""ctypes binding of libmv3d_b200.so (the C ABI in include/mv3d_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.  PyTorch tensors
are used only as device-memory containers (`tensor.data_ptr()`) and for the current stream.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmv3d_b200.so")

c_void_p, c_int, c_float, c_double, c_size_t = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t


class ProposalParams(C.Structure):
    _fields_ = [("Hf", c_int), ("Wf", c_int), ("A", c_int),
                ("xn", c_double), ("yn", c_double), ("x_min", c_double), ("y_min", c_double), ("res", c_double),
                ("im_h", c_float), ("im_w", c_float), ("im_scale", c_float),
                ("img_h", c_float), ("img_w", c_float), ("min_size", c_float),
                ("pre_nms_top_n", c_int), ("post_nms_top_n", c_int),
                ("nms_thresh", c_double), ("nms_rule_ge", c_int), ("batch_index", c_float), ("d_proj", c_void_p),
                ("ld_prob", c_int), ("ld_deltas", c_int)]


class RoiView(C.Structure):
    _fields_ = [("d_data", c_void_p), ("d_rois", c_void_p), ("height", c_int), ("width", c_int),
                ("spatial_scale", c_float), ("d_top", c_void_p), ("d_argmax", c_void_p),
                ("d_top_hi", c_void_p), ("d_top_lo", c_void_p), ("source", c_int), ("d_rois_out", c_void_p),
                ("d_pad_hi", c_void_p), ("d_pad_lo", c_void_p), ("pad_fmt", c_int), ("pad_c", c_int),
                ("top_fmt", c_int)]


ROI_GIVEN, ROI_BEV, ROI_IMG, ROI_FV = 0, 1, 2, 3   # MV3D_ROI_* of include/mv3d_b200.h


class RoiProjection(C.Structure):
    _fields_ = [("xn", c_double), ("yn", c_double), ("x_min", c_double), ("y_min", c_double), ("res", c_double),
                ("im_h", c_float), ("im_w", c_float), ("h_proj", c_float * 12), ("d_proj", c_void_p),
                ("fv_h", c_int), ("fv_w", c_int), ("fv_theta_min", c_double), ("fv_dtheta", c_double),
                ("fv_phi_max", c_double), ("fv_dphi", c_double)]


class GemmDesc(C.Structure):
    _fields_ = [("M", c_int), ("N", c_int), ("Cin", c_int), ("taps", c_int), ("Hp", c_int), ("Wp", c_int),
                ("passes", c_int),
                ("d_a_hi", c_void_p), ("d_a_lo", c_void_p), ("d_w_hi", c_void_p), ("d_w_lo", c_void_p),
                ("d_bias", c_void_p), ("relu", c_int),
                ("d_out_hi", c_void_p), ("d_out_lo", c_void_p), ("ld_out", c_int),
                ("d_out_f32", c_void_p), ("ld_f32", c_int), ("f32_dense", c_int), ("split_k", c_int),
                ("d_mask_hi", c_void_p), ("ld_mask", c_int), ("mask_scale", c_float),
                ("d_addend_f32", c_void_p), ("ld_addend", c_int), ("out_fmt", c_int), ("softmax_cols", c_int), ("pool", c_int),
                ("cin_valid", c_int)]


class WgradDesc(C.Structure):
    _fields_ = [("P", c_int), ("Cx", c_int), ("Cg", c_int), ("cin", c_int), ("cout", c_int), ("taps", c_int),
                ("Wp", c_int), ("passes", c_int),
                ("d_x_hi", c_void_p), ("d_x_lo", c_void_p), ("d_g_hi", c_void_p), ("d_g_lo", c_void_p),
                ("d_dw", c_void_p), ("ld_dw", c_int), ("accumulate", c_int), ("split_rows", c_int),
                ("tap_window", c_int)]


# name -> (restype, argtypes); every symbol include/mv3d_b200.h declares
SIGNATURES = {
    "mv3d_version": (c_int, []),
    "mv3d_status_string": (C.c_char_p, [c_int]),
    "mv3d_last_cuda_error": (c_int, []),
    "mv3d_last_cuda_error_string": (C.c_char_p, []),
    "mv3d_bev_raster_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "mv3d_bev_raster": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                c_float, c_float, c_float, c_float, c_float, c_float, c_int, c_int, c_void_p,
                                c_size_t, c_void_p]),
    "mv3d_bev_raster_pad": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                    c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_float, c_float, c_int,
                                    c_int, c_void_p, c_size_t, c_void_p]),
    "mv3d_bev_raster_pad_fmt": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_float, c_float, c_int,
                                        c_int, c_void_p, c_size_t, c_int, c_void_p]),
    "_nms": (None, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_int]),
    "mv3d_nms_workspace_bytes": (c_size_t, [c_int]),
    "mv3d_nms": (c_int, [c_void_p, c_int, c_int, c_void_p, c_double, c_int, c_int, c_void_p, c_void_p, c_void_p,
                         c_size_t, c_void_p]),
    "mv3d_proposal_workspace_bytes": (c_size_t, [C.POINTER(ProposalParams)]),
    "mv3d_proposal_layer_3d": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.POINTER(ProposalParams), c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                       c_void_p]),
    "mv3d_proposal_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.POINTER(ProposalParams), c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mv3d_roi_pool_forward": (c_int, [c_void_p, c_float, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p]),
    "mv3d_roi_pool_backward": (c_int, [c_void_p, c_float, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                       c_void_p, c_void_p, c_void_p]),
    "mv3d_roi_pool_multiview": (c_int, [C.POINTER(RoiView), c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "mv3d_roi_pool_fused": (c_int, [C.POINTER(RoiView), c_int, c_void_p, C.POINTER(RoiProjection), c_int, c_void_p,
                                    c_int, c_int, c_int, c_void_p]),
    "mv3d_conv_gemm": (c_int, [C.POINTER(GemmDesc), c_void_p]),
    "mv3d_gemm_set_pair_mode": (c_int, [c_int]),
    "mv3d_gemm_set_stamps": (c_int, [c_void_p]),
    "mv3d_pack_weights_fmt": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "mv3d_pad_nhwc_fmt": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "mv3d_unpad_nhwc_fmt": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "mv3d_maxpool2x2_pad_fmt": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                        c_void_p]),
    "mv3d_conv_wgrad": (c_int, [C.POINTER(WgradDesc), c_void_p]),
    "mv3d_pack_weights": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mv3d_pad_nhwc": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mv3d_im2col3x3_pad": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mv3d_conv3x3_small_cin": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                       c_void_p, c_int, c_int, c_void_p]),
    "mv3d_unpad_nhwc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mv3d_maxpool2x2_pad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mv3d_softmax_pairs": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "mv3d_maxpool2x2_bwd_pad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                        c_void_p, c_void_p]),
    "mv3d_bias_grad": (c_int, [c_void_p, c_void_p, C.c_longlong, c_int, c_int, c_void_p, c_void_p]),
    "mv3d_pack_weights_dgrad": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mv3d_pad_nhwc_masked": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    "mv3d_dropout": (c_int, [c_void_p, c_void_p, C.c_longlong, c_int, c_int, c_float, C.c_ulonglong, c_void_p]),
    "mv3d_rpn_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                              c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mv3d_rcnn_loss": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int,
                               c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mv3d_adam": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.c_longlong, c_float, c_float, c_float, c_float,
                          c_int, c_float, c_void_p]),
    "mv3d_anchor_targets": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_float, c_float, c_double,
                                    c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mv3d_roi_overlaps": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "mv3d_proposal_targets": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                      c_int, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p]),
    "mv3d_fv_raster_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mv3d_fv_raster": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_double, c_double, c_double, c_double, c_void_p,
                               c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "mv3d_rois_to_fv": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_double, c_double, c_double, c_double,
                                c_void_p, c_void_p]),
    "mv3d_bias_act": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                              c_int, c_void_p]),
}

_lib = None


class Mv3dError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the CUDA library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Mv3dError("%s not built -- run `python -m mv3d_tf_b200.build` (there is no CPU fallback)" % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if not hasattr(handle, name) and os.environ.get("MV3D_DEV_PARTIAL_LIB") == "1":
                continue  # development only: library built from a subset of the sources
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


# kernels launched by one successful call of each entry point (memsets not counted) -- bench.py's gpu_launches
KERNELS_PER_CALL = {"mv3d_bev_raster": 4, "mv3d_bev_raster_pad": 4, "mv3d_bev_raster_pad_fmt": 4, "mv3d_nms": 2, "mv3d_proposal_layer_3d": 7, "mv3d_proposal_decode": 1,
                    "mv3d_roi_pool_forward": 1, "mv3d_roi_pool_backward": 1, "mv3d_roi_pool_multiview": 1,
                    "mv3d_conv_gemm": 1, "mv3d_conv_wgrad": 1, "mv3d_pack_weights": 1, "mv3d_pad_nhwc": 1, "mv3d_im2col3x3_pad": 1, "mv3d_unpad_nhwc": 1,
                    "mv3d_maxpool2x2_pad": 1, "mv3d_pack_weights_fmt": 1, "mv3d_pad_nhwc_fmt": 1, "mv3d_unpad_nhwc_fmt": 1,
                    "mv3d_maxpool2x2_pad_fmt": 1, "mv3d_softmax_pairs": 1, "mv3d_bias_act": 1,
                    "mv3d_maxpool2x2_bwd_pad": 1, "mv3d_bias_grad": 1, "mv3d_pack_weights_dgrad": 1,
                    "mv3d_pad_nhwc_masked": 1, "mv3d_dropout": 1, "mv3d_rpn_loss": 1, "mv3d_rcnn_loss": 1,
                    "mv3d_adam": 1, "mv3d_fv_raster": 2, "mv3d_rois_to_fv": 1, "mv3d_anchor_targets": 2, "mv3d_roi_overlaps": 1, "mv3d_proposal_targets": 1}
_launches = 0


def reset_launch_count() -> None:
    global _launches
    _launches = 0


def launch_count() -> int:
    return _launches


def check(status: int, what: str = "") -> None:
    global _launches
    _launches += KERNELS_PER_CALL.get(what, 0)
    if status != 0:
        L = lib()
        msg = L.mv3d_status_string(status).decode()
        if status == -3:
            msg += " [%s]" % L.mv3d_last_cuda_error_string().decode()
        raise Mv3dError("%s failed: %s" % (what or "mv3d call", msg))


def ptr(t) -> int:
    """Device/host address of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream
