"""Host-side constants of the BEV <-> LiDAR <-> image geometry (lib/utils/transform.py:3-20,81-111,369-386).
Only what must be computed once per network / per frame on the host lives here (the anchor table and
the 3x4 projection matrix); the per-box work runs in csrc/proposal.cu."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

LIDAR_HEIGHT = 1.73
CAR_HEIGHT = 1.56


@dataclass(frozen=True)
class BevGeometry:
    """transform.py:3-11.  Defaults are the reference's only extent (601x601 BEV)."""

    x_min: float = 0
    x_max: float = 60
    y_min: float = -30
    y_max: float = 30
    res: float = 0.1

    @property
    def xn(self) -> int:
        return int((self.x_max - self.x_min) // self.res) + 1

    @property
    def yn(self) -> int:
        return int((self.y_max - self.y_min) // self.res) + 1

    @property
    def is_reference(self) -> bool:
        return self == BevGeometry()


REF_GEOMETRY = BevGeometry()
CFG_GEOMETRY = BevGeometry(0, 70, -40, 40, 0.1)


def bv_anchor_to_lidar(anchors: np.ndarray, geom: BevGeometry = REF_GEOMETRY) -> np.ndarray:
    """transform.py:89-111 -> (N,6) float64 [x,y,z,l,w,h].  On the reference grid the bv-x axis is scaled
    with Xn and bv-y with Yn exactly as the reference does (both 600 there); other grids pair each axis
    with its own extent."""
    a = anchors.astype(np.int64, copy=False)
    n = a.shape[0]
    lengths = (a[:, 3] - a[:, 1]).reshape(n, 1) * geom.res
    widths = (a[:, 2] - a[:, 0]).reshape(n, 1) * geom.res
    ctr_bx = ((a[:, 0] + a[:, 2]) / 2.).reshape(n, 1)
    ctr_by = ((a[:, 1] + a[:, 3]) / 2.).reshape(n, 1)
    nx, ny = (geom.xn, geom.yn) if geom.is_reference else (geom.yn, geom.xn)
    y = nx * geom.res - (ctr_bx + 0.5) * geom.res + geom.y_min
    x = ny * geom.res - (ctr_by + 0.5) * geom.res + geom.x_min
    z = np.ones((n, 1), dtype=np.float32) * -(LIDAR_HEIGHT - CAR_HEIGHT / 2.)
    h = np.ones((n, 1), dtype=np.float32) * CAR_HEIGHT
    return np.hstack((x, y, z, lengths, widths, h))


def projection_matrix(calib: np.ndarray) -> np.ndarray:
    """(P2 . R0[4x3]) . Tr of transform.py:371-384 from the (4,12) calib blob (rows P2, P3, R0, Tr;
    kitti_mv3d.py:63-75), in the dtype calib arrives in (float32 at the layer boundary) -> (3,4) float32."""
    calib = np.asarray(calib)
    tr, r0, p2 = calib[3].reshape(3, 4), calib[2].reshape(4, 3), calib[0].reshape(3, 4)
    return np.ascontiguousarray(np.dot(np.dot(p2, r0), tr), dtype=np.float32)


@dataclass(frozen=True)
class FvGeometry:
    """Cylindrical front-view map (no reference counterpart; specification: DESIGN.md, section 'Front view'):
    H rows over elevation [phi_min, phi_max] degrees (row 0 = top), W columns over azimuth [theta_min, theta_max)."""

    H: int = 64
    W: int = 512
    theta_min: float = -45.0
    theta_max: float = 45.0
    phi_min: float = -24.9
    phi_max: float = 2.0

    @property
    def dtheta(self) -> float:
        return float(np.deg2rad(self.theta_max - self.theta_min) / self.W)

    @property
    def dphi(self) -> float:
        return float(np.deg2rad(self.phi_max - self.phi_min) / self.H)

    def c_args(self):
        """(H, W, theta_min_rad, dtheta, phi_max_rad, dphi) in the order the C ABI takes them."""
        return (self.H, self.W, float(np.deg2rad(self.theta_min)), self.dtheta, float(np.deg2rad(self.phi_max)), self.dphi)


FV_GEOMETRY = FvGeometry()
