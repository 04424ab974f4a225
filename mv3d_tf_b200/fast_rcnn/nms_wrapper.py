"""Dispatcher with the reference's signature (lib/fast_rcnn/nms_wrapper.py:13-21)."""
from .config import cfg
from ..nms.gpu_nms import cpu_nms, gpu_nms


def nms(dets, thresh, force_cpu=False):
    """USE_GPU_NMS -> `>` rule of nms_kernel.cu; otherwise the `>=` rule of cpu_nms.pyx.
    Both run on the GPU here; the flag only selects which reference rule is reproduced."""
    if dets.shape[0] == 0:
        return []
    if cfg.USE_GPU_NMS and not force_cpu:
        return gpu_nms(dets, thresh, device_id=cfg.GPU_ID)
    return cpu_nms(dets, thresh)
