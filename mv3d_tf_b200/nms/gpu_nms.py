"""NMS entry points with the reference's signatures.
  gpu_nms(dets, thresh, device_id=0)  -- lib/nms/gpu_nms.pyx:16-31 (host argsort, then the C ABI `_nms`)
  cpu_nms(dets, thresh)               -- lib/nms/cpu_nms.pyx:17-68 semantics (`>=` in double) on the GPU
Both return a python list of kept indices into `dets`, score-descending."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .._lib import check, current_stream, lib, ptr


def _order_desc(scores: np.ndarray) -> np.ndarray:
    # argsort()[::-1] with ties resolved higher-index-first (stable ascending sort, reversed)
    return np.argsort(scores, kind="stable")[::-1]


def gpu_nms(dets: np.ndarray, thresh: float, device_id: int = 0):
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    n, dim = dets.shape
    if n == 0:
        return []
    order = _order_desc(dets[:, 4])
    sorted_dets = np.ascontiguousarray(dets[order, :])
    keep = np.zeros(n, dtype=np.int32)
    num = C.c_int(0)
    lib()._nms(keep.ctypes.data, C.addressof(num), sorted_dets.ctypes.data, n, dim, C.c_float(thresh), device_id)
    return list(order[keep[: num.value]])


def nms_device(boxes: torch.Tensor, thresh: float, rule_ge: bool = True, max_keep: int = 0):
    """Device form: `boxes` (n, >=4) float32 CUDA tensor ALREADY sorted by score descending.
    Returns (keep int32 tensor of capacity n, count int32 tensor[1]) without synchronising."""
    assert boxes.is_cuda and boxes.dtype == torch.float32 and boxes.is_contiguous()
    n = boxes.shape[0]
    L = lib()
    keep = torch.empty(max(n, 1), dtype=torch.int32, device=boxes.device)
    num = torch.zeros(1, dtype=torch.int32, device=boxes.device)
    ws = torch.empty(L.mv3d_nms_workspace_bytes(n), dtype=torch.uint8, device=boxes.device)
    check(L.mv3d_nms(ptr(boxes), n, boxes.shape[1], None, float(thresh), int(rule_ge), int(max_keep), ptr(keep),
                     ptr(num), ptr(ws), ws.numel(), current_stream()), "mv3d_nms")
    return keep, num


def cpu_nms(dets: np.ndarray, thresh: float):
    """cpu_nms.pyx semantics (suppress when (double)ovr >= thresh), computed on the GPU."""
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.shape[0] == 0:
        return []
    order = _order_desc(dets[:, 4])
    boxes = torch.from_numpy(np.ascontiguousarray(dets[order, :4])).cuda()
    keep, num = nms_device(boxes, thresh, rule_ge=True)
    k = keep[: int(num.item())].cpu().numpy()
    return list(order[k])
