"""Label geometry of the KITTI feed (host side, numpy, once per label at roidb-build time):
lib/utils/transform.py:13-20,113-142,172-187,441-465,502-525 -- camera-frame box -> 8 corners -> LiDAR frame ->
LiDAR (x,y,z,l,w,h) -> bird's-eye-view box.  Arithmetic and dtypes as in the reference (float64 intermediates, numpy's
float floor-division for the BEV cell index, float32 result)."""
import numpy as np

from .transform import REF_GEOMETRY, BevGeometry


def _lidar_to_bv_coord(x, y, geom: BevGeometry = REF_GEOMETRY):
    xx = geom.yn - (y - geom.y_min) // geom.res
    yy = geom.xn - (x - geom.x_min) // geom.res
    return xx, yy


def lidar_3d_to_bv(rois_3d, geom: BevGeometry = REF_GEOMETRY):
    """(x,y,z,l,w,h) LiDAR box(es) -> (x1,y1,x2,y2) BEV, float32 (transform.py:113-142)."""
    rois_3d = np.asarray(rois_3d)
    if rois_3d.ndim == 1:
        rois = np.zeros(4)
        rois[0] = rois_3d[0] + rois_3d[3] * 0.5
        rois[1] = rois_3d[1] + rois_3d[4] * 0.5
        rois[2] = rois_3d[0] - rois_3d[3] * 0.5
        rois[3] = rois_3d[1] - rois_3d[4] * 0.5
        rois[0], rois[1] = _lidar_to_bv_coord(rois[0], rois[1], geom)
        rois[2], rois[3] = _lidar_to_bv_coord(rois[2], rois[3], geom)
    else:
        rois = np.zeros((rois_3d.shape[0], 4))
        rois[:, 0] = rois_3d[:, 0] + rois_3d[:, 3] * 0.5
        rois[:, 1] = rois_3d[:, 1] + rois_3d[:, 4] * 0.5
        rois[:, 2] = rois_3d[:, 0] - rois_3d[:, 3] * 0.5
        rois[:, 3] = rois_3d[:, 1] - rois_3d[:, 4] * 0.5
        rois[:, 0], rois[:, 1] = _lidar_to_bv_coord(rois[:, 0], rois[:, 1], geom)
        rois[:, 2], rois[:, 3] = _lidar_to_bv_coord(rois[:, 2], rois[:, 3], geom)
    return rois.astype(np.float32)


def lidar_cnr_to_3d(corners, lwh):
    """24 corner coordinates (x0..x7,y0..y7,z0..z7) + (l,w,h) -> (cx,cy,cz,l,w,h) float64 (transform.py:172-187)."""
    corners = np.asarray(corners)
    if corners.shape[0] == 24 and corners.ndim == 1:
        boxes_3d = np.zeros(6)
        boxes_3d[:3] = corners.reshape((3, 8)).mean(1)
        boxes_3d[3:] = lwh
    else:
        boxes_3d = np.zeros((corners.shape[0], 6))
        boxes_3d[:, :3] = corners.reshape((-1, 3, 8)).mean(2)
        boxes_3d[:, 3:] = lwh
    return boxes_3d


def computeCorners3D(Boxex3D, ry):
    """Camera-frame box (x,y,z,l,w,h) + yaw -> (3,8) corners (transform.py:441-465)."""
    R = np.array([[np.cos(ry), 0, np.sin(ry)], [0, 1, 0], [-np.sin(ry), 0, np.cos(ry)]]).reshape((3, 3))
    l, w, h = Boxex3D[3:6]
    x, y, z = Boxex3D[0:3]
    x_corners = np.array([l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2, -l / 2, -l / 2])
    y_corners = np.array([0, 0, 0, 0, -h, -h, -h, -h])
    z_corners = np.array([w / 2, -w / 2, -w / 2, w / 2, w / 2, -w / 2, -w / 2, w / 2])
    corners_3D = np.dot(R, np.vstack((x_corners, y_corners, z_corners)))
    corners_3D[0, :] = corners_3D[0, :] + x
    corners_3D[1, :] = corners_3D[1, :] + y
    corners_3D[2, :] = corners_3D[2, :] + z
    return corners_3D


def camera_to_lidar_cnr(pts_3D, P):
    """(3,8) camera corners -> (1,24) LiDAR corners with the inverse rotation of Tr_velo_to_cam and the reference's
    permuted translation (transform.py:502-525: T = (-P[1,3], -P[2,3], P[0,3]); the homogeneous row is zeros, so T
    never contributes -- reproduced as is)."""
    if pts_3D.shape[1] == 24:
        pts_3D = pts_3D.reshape((3, 8))
    pts_3D = np.vstack((pts_3D, np.zeros(8)))
    assert pts_3D.shape == (4, 8)
    R = np.linalg.inv(P[:, :3])
    T = np.zeros((3, 1))
    T[0] = -P[1, 3]
    T[1] = -P[2, 3]
    T[2] = P[0, 3]
    lidar_corners = np.dot(np.hstack((R, T)), pts_3D)[:3, :]
    return lidar_corners.reshape(-1, 24)
