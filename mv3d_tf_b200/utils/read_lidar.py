"""LiDAR -> BEV raster with the reference's call signature (tools/read_lidar.py:10-115 ==
lib/utils/read_lidar.py:10-115); the work runs in csrc/bev_raster.cu."""
from __future__ import annotations

import numpy as np
import torch

from .._lib import check, current_stream, lib, ptr


def raster_geometry(res, zres, side_range, fwd_range, height_range):
    """The host scalars of read_lidar.py:49-53,80,102-103 (same Python/numpy expressions, so the same roundings)."""
    x_max = int((side_range[1] - side_range[0]) / res)
    y_max = int((fwd_range[1] - fwd_range[0]) / res)
    z_max = int((height_range[1] - height_range[0]) / zres)
    lows = np.arange(height_range[0], height_range[1], zres)
    lo = np.ascontiguousarray(lows, dtype=np.float64)
    hi = np.ascontiguousarray([h + zres for h in lows], dtype=np.float64)
    return dict(H=y_max + 1, W=x_max + 1, C=z_max + 1, nslices=int(lo.shape[0]), lo=lo, hi=hi,
                xoff=int(np.floor(side_range[0] / res)), yoff=int(np.floor(fwd_range[1] / res)))


class BevRasterizer:
    """Reusable rasteriser for one grid configuration: owns the workspace, returns a device tensor."""

    def __init__(self, res=0.1, zres=0.3, side_range=(-30., 30.), fwd_range=(0., 60.), height_range=(-2., 0.4),
                 device="cuda"):
        self.args = (res, zres, tuple(side_range), tuple(fwd_range), tuple(height_range))
        self.g = raster_geometry(*self.args)
        if self.g["nslices"] > self.g["C"]:
            raise ValueError("more height slices than channels: the reference would raise IndexError")
        self.device = torch.device(device)
        self._ws = None
        self._ws_points = -1

    @property
    def shape(self):
        return (self.g["H"], self.g["W"], self.g["C"])

    def __call__(self, points: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        assert points.is_cuda and points.dtype == torch.float32 and points.dim() == 2 and points.shape[1] >= 4
        points = points.contiguous()
        n = points.shape[0]
        g = self.g
        L = lib()
        self._workspace(n, points.device)
        if out is None:
            out = torch.empty(self.shape, dtype=torch.float32, device=points.device)
        res, zres, side, fwd, hr = self.args
        check(L.mv3d_bev_raster(ptr(points), n, points.shape[1], ptr(out), g["H"], g["W"], g["C"], g["nslices"],
                                ptr(g["lo"]), ptr(g["hi"]), res, fwd[0], fwd[1], side[0], side[1], hr[0], g["xoff"],
                                g["yoff"], ptr(self._ws), self._ws.numel(), current_stream()), "mv3d_bev_raster")
        return out


    def _workspace(self, n, device):
        g = self.g
        if n > self._ws_points:
            nbytes = lib().mv3d_bev_raster_workspace_bytes(max(n, 1), g["H"], g["W"], g["nslices"])
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._ws_points = n
        return self._ws

    def to_pad(self, clouds, precise: bool = True, fmt: int = 0):
        """Rasterise one cloud per frame straight into the conv trunk's input layout (kernels.PadAct, zero halo) -- the
        float32 map is never materialised.  fmt: kernels.FMT_BF16X2 (bf16 hi/lo) or kernels.FMT_F16E5 (the mixed-mode
        operand format; channels padded to a multiple of 64)."""
        from ..kernels import BF16, FMT_F16E5, PadAct, pad_channels, round_up

        if isinstance(clouds, torch.Tensor):
            clouds = [clouds]
        g = self.g
        cp = round_up(g["C"], 64) if fmt == FMT_F16E5 else pad_channels(g["C"])
        dev = clouds[0].device
        hi = torch.empty((len(clouds), g["H"] + 1, g["W"] + 1, cp), dtype=BF16, device=dev)
        lo = torch.empty_like(hi) if (precise or fmt == FMT_F16E5) else None
        res, zres, side, fwd, hr = self.args
        for b, pts in enumerate(clouds):
            assert pts.is_cuda and pts.dtype == torch.float32 and pts.dim() == 2 and pts.shape[1] >= 4
            pts = pts.contiguous()
            ws = self._workspace(pts.shape[0], dev)
            check(lib().mv3d_bev_raster_pad_fmt(ptr(pts), pts.shape[0], pts.shape[1], ptr(hi[b]),
                                                ptr(lo[b]) if lo is not None else None, cp, g["H"], g["W"], g["C"],
                                                g["nslices"], ptr(g["lo"]), ptr(g["hi"]), res, fwd[0], fwd[1], side[0],
                                                side[1], hr[0], g["xoff"], g["yoff"], ptr(ws), ws.numel(), fmt,
                                                current_stream()), "mv3d_bev_raster_pad_fmt")
        return PadAct(hi, lo, len(clouds), g["H"], g["W"], g["C"], fmt)


def point_cloud_2_top(points, res=0.1, zres=0.3, side_range=(-10., 10.), fwd_range=(-10., 10.),
                      height_range=(-2., 2.)):
    """Drop-in for the reference function: numpy (N,>=4) in, numpy (H,W,C) float32 out."""
    pts = torch.from_numpy(np.ascontiguousarray(points[:, :4], dtype=np.float32)).cuda()
    r = BevRasterizer(res, zres, side_range, fwd_range, height_range, device=pts.device)
    return r(pts).cpu().numpy()


class FvRasterizer:
    """LiDAR -> cylindrical front-view map (H,W,3) [z, range, reflectance] (csrc/front_view.cu).  The reference has no
    front view; see utils.transform.FvGeometry."""

    def __init__(self, geom=None, device="cuda"):
        from .transform import FV_GEOMETRY

        self.g = geom or FV_GEOMETRY
        self.device = torch.device(device)
        self._ws = torch.empty(lib().mv3d_fv_raster_workspace_bytes(self.g.H, self.g.W), dtype=torch.uint8,
                               device=self.device)

    @property
    def shape(self):
        return (self.g.H, self.g.W, 3)

    def _call(self, pts, top, hi, lo, c_pad):
        assert pts.is_cuda and pts.dtype == torch.float32 and pts.dim() == 2 and pts.shape[1] >= 4
        pts = pts.contiguous()
        H, W, t0, dt, p1, dp = self.g.c_args()
        check(lib().mv3d_fv_raster(ptr(pts), pts.shape[0], pts.shape[1], H, W, t0, dt, p1, dp, ptr(top), ptr(hi),
                                   ptr(lo), c_pad, ptr(self._ws), self._ws.numel(), current_stream()), "mv3d_fv_raster")

    def __call__(self, points: torch.Tensor) -> torch.Tensor:
        out = torch.empty(self.shape, dtype=torch.float32, device=points.device)
        self._call(points, out, None, None, 0)
        return out

    def to_pad(self, clouds, precise: bool = True):
        """One cloud per frame -> the trunk's PAD input (bf16 hi/lo, zero halo, 16 channels)."""
        from ..kernels import BF16, PadAct, pad_channels

        if isinstance(clouds, torch.Tensor):
            clouds = [clouds]
        cp = pad_channels(3)
        dev = clouds[0].device
        hi = torch.empty((len(clouds), self.g.H + 1, self.g.W + 1, cp), dtype=BF16, device=dev)
        lo = torch.empty_like(hi) if precise else None
        for b, pts in enumerate(clouds):
            self._call(pts, None, hi[b], lo[b] if lo is not None else None, cp)
        return PadAct(hi, lo, len(clouds), self.g.H, self.g.W, 3)


def point_cloud_2_front(points, geom=None):
    """numpy (N,>=4) in, numpy (H,W,3) float32 out."""
    pts = torch.from_numpy(np.ascontiguousarray(points[:, :4], dtype=np.float32)).cuda()
    return FvRasterizer(geom, device=pts.device)(pts).cpu().numpy()
