"""KITTI object-detection set in the form MV3D consumes (lib/datasets/kitti_mv3d.py:20-352 + the parts of
lib/datasets/imdb.py it inherits): image / BEV paths, the 4x12 calib blob, ground-truth boxes in image, camera,
LiDAR, corner and bird's-eye-view form, the KITTI results writer.

Same method names, dictionary keys, dtypes and arithmetic as the reference (label geometry goes through
utils/transform.py's computeCorners3D -> camera_to_lidar_cnr -> lidar_cnr_to_3d -> lidar_3d_to_bv exactly as
kitti_mv3d.py:263-271 chains them).  Two things are new, both optional: `frame_at(i)` for the inference loop, and
the online rasteriser -- when `<split>/lidar_bv/<index>.npy` (the reference's offline raster, tools/read_lidar.py:125-135)
is missing but `<split>/velodyne/<index>.bin` exists, the BEV map is rasterised on the GPU from the raw cloud.
"""
from __future__ import annotations

import os
import pickle
import time

import numpy as np

from . import ROOT_DIR
from ..utils.kitti_geometry import camera_to_lidar_cnr, computeCorners3D, lidar_3d_to_bv, lidar_cnr_to_3d


class imdb(object):
    """The slice of lib/datasets/imdb.py:17-103 the MV3D path uses."""

    def __init__(self, name):
        self._name = name
        self._classes = []
        self._image_index = []
        self._roidb = None
        self._roidb_handler = self.default_roidb
        self.config = {}

    @property
    def name(self):
        return self._name

    @property
    def num_classes(self):
        return len(self._classes)

    @property
    def classes(self):
        return self._classes

    @property
    def image_index(self):
        return self._image_index

    @property
    def roidb_handler(self):
        return self._roidb_handler

    @roidb_handler.setter
    def roidb_handler(self, val):
        self._roidb_handler = val

    @property
    def roidb(self):
        # imdb.py:62-72: computed once by the handler
        if self._roidb is not None:
            return self._roidb
        self._roidb = self.roidb_handler()
        return self._roidb

    @property
    def cache_path(self):
        cache_path = os.path.abspath(os.path.join(self._cache_root(), 'cache'))
        if not os.path.exists(cache_path):
            os.makedirs(cache_path)
        return cache_path

    def _cache_root(self):
        return os.path.join(ROOT_DIR, 'data')

    @property
    def num_images(self):
        return len(self.image_index)

    def default_roidb(self):
        raise NotImplementedError

    def competition_mode(self, on):
        pass


class kitti_mv3d(imdb):
    def __init__(self, image_set, kitti_path=None, use_cache=True):
        imdb.__init__(self, image_set)   # kitti_mv3d.py:22: the imdb name is the split
        self._image_set = image_set
        self._kitti_path = self._get_default_path() if kitti_path is None else kitti_path
        self._data_path = os.path.join(self._kitti_path, 'object')
        self._classes = ('__background__', 'Car')
        self._class_to_ind = dict(zip(self.classes, range(self.num_classes)))
        self._image_ext = '.png'
        self._lidar_ext = '.npy'
        self._subset = 'car'
        self._use_cache = use_cache
        assert os.path.exists(self._kitti_path), 'KITTI path does not exist: {}'.format(self._kitti_path)
        assert os.path.exists(self._data_path), 'Path does not exist: {}'.format(self._data_path)
        self._image_index = self._load_image_set_index()
        self._roidb_handler = self.gt_roidb
        self.config = {'top_k': 100000}
        self._rasterizer = None
        # the raster the reference writes offline (tools/read_lidar.py:121-123)
        self.raster_args = dict(res=0.1, zres=0.3, side_range=(-30., 30.), fwd_range=(0., 60.), height_range=(-2., 0.4))

    def _cache_root(self):
        return self._kitti_path

    def _prefix(self, what):
        return ('testing/' if self._image_set == 'test' else 'training/') + what

    # ------------------------------------------------------------------ paths (kitti_mv3d.py:50-104)
    def image_path_at(self, i):
        return self.image_path_from_index(self.image_index[i])

    def lidar_path_at(self, i):
        return self.lidar_path_from_index(self.image_index[i])

    def image_path_from_index(self, index):
        image_path = os.path.join(self._data_path, self._prefix('image_2'), index + self._image_ext)
        assert os.path.exists(image_path), 'Path does not exist: {}'.format(image_path)
        return image_path

    def lidar_path_from_index(self, index):
        """The offline BEV raster; if it is absent the raw cloud's path is returned instead (rasterised online)."""
        lidar_bv_path = os.path.join(self._data_path, self._prefix('lidar_bv'), index + self._lidar_ext)
        if os.path.exists(lidar_bv_path):
            return lidar_bv_path
        velo = self.velodyne_path_from_index(index)
        assert os.path.exists(velo), 'Path does not exist: {}'.format(lidar_bv_path)
        return velo

    def velodyne_path_from_index(self, index):
        return os.path.join(self._data_path, self._prefix('velodyne'), index + '.bin')

    def calib_at(self, i):
        """(4,12) float64 rows P2, P3, R0 (9 values, rest 0), Tr_velo_to_cam (kitti_mv3d.py:63-75).  As in the
        reference the file is chosen by the POSITION i (`str(i).zfill(6)`), not by image_index[i]."""
        index = str(i).zfill(6)
        calib_ori = self._load_kitti_calib(index)
        calib = np.zeros((4, 12))
        calib[0, :] = calib_ori['P2'].reshape(12)
        calib[1, :] = calib_ori['P3'].reshape(12)
        calib[2, :9] = calib_ori['R0'].reshape(9)
        calib[3, :] = calib_ori['Tr_velo2cam'].reshape(12)
        return calib

    def _load_image_set_index(self):
        image_set_file = os.path.join(self._kitti_path, 'ImageSets', self._image_set + '.txt')
        assert os.path.exists(image_set_file), 'Path does not exist: {}'.format(image_set_file)
        with open(image_set_file) as f:
            image_index = [x.rstrip('\n') for x in f.readlines()]
        print('image sets length: ', len(image_index))
        return image_index

    def _get_default_path(self):
        return os.path.join(ROOT_DIR, 'data', 'KITTI')

    # ------------------------------------------------------------------ ground truth (kitti_mv3d.py:128-306)
    def gt_roidb(self):
        cache_file = os.path.join(self.cache_path, self.name + '_gt_roidb.pkl')
        if self._use_cache and os.path.exists(cache_file):
            with open(cache_file, 'rb') as fid:
                roidb = pickle.load(fid)
            print('{} gt roidb loaded from {}'.format(self.name, cache_file))
            return roidb
        gt_roidb = [self._load_kitti_annotation(index) for index in self.image_index]
        if self._use_cache:
            with open(cache_file, 'wb') as fid:
                pickle.dump(gt_roidb, fid, pickle.HIGHEST_PROTOCOL)
            print('wrote gt roidb to {}'.format(cache_file))
        return gt_roidb

    def _load_kitti_calib(self, index):
        calib_dir = os.path.join(self._data_path, self._prefix('calib'), index + '.txt')
        with open(calib_dir) as fi:
            lines = fi.readlines()
        row = lambda k: np.array(lines[k].strip().split(' ')[1:], dtype=np.float32)
        return {'P2': row(2).reshape(3, 4), 'P3': row(3).reshape(3, 4), 'R0': row(4).reshape(3, 3),
                'Tr_velo2cam': row(5).reshape(3, 4)}

    def _load_kitti_annotation(self, index):
        filename = os.path.join(self._data_path, 'training/label_2', index + '.txt')
        Tr = self._load_kitti_calib(index)['Tr_velo2cam']
        with open(filename, 'r') as f:
            lines = f.readlines()
        num_objs = len(lines)
        translation = np.zeros((num_objs, 3), dtype=np.float32)
        rys = np.zeros((num_objs), dtype=np.float32)
        lwh = np.zeros((num_objs, 3), dtype=np.float32)
        boxes = np.zeros((num_objs, 4), dtype=np.float32)
        boxes_bv = np.zeros((num_objs, 4), dtype=np.float32)
        boxes3D = np.zeros((num_objs, 6), dtype=np.float32)
        boxes3D_lidar = np.zeros((num_objs, 6), dtype=np.float32)
        boxes3D_cam_cnr = np.zeros((num_objs, 24), dtype=np.float32)
        boxes3D_corners = np.zeros((num_objs, 24), dtype=np.float32)
        alphas = np.zeros((num_objs), dtype=np.float32)
        gt_classes = np.zeros((num_objs), dtype=np.int32)
        overlaps = np.zeros((num_objs, self.num_classes), dtype=np.float32)
        ix = -1
        for line in lines:
            obj = line.strip().split(' ')
            cls = self._class_to_ind.get(obj[0].strip())
            if cls is None:      # other KITTI classes are skipped (kitti_mv3d.py:230-234)
                continue
            ix += 1
            alpha = float(obj[3])
            x1, y1, x2, y2 = float(obj[4]), float(obj[5]), float(obj[6]), float(obj[7])
            h, w, l = float(obj[8]), float(obj[9]), float(obj[10])
            tx, ty, tz = float(obj[11]), float(obj[12]), float(obj[13])
            ry = float(obj[14])
            rys[ix] = ry
            lwh[ix, :] = [l, w, h]
            alphas[ix] = alpha
            translation[ix, :] = [tx, ty, tz]
            boxes[ix, :] = [x1, y1, x2, y2]
            boxes3D[ix, :] = [tx, ty, tz, l, w, h]
            cam_cnr = computeCorners3D(boxes3D[ix, :], ry)                 # 8 corners, camera frame
            boxes3D_cam_cnr[ix, :] = cam_cnr.reshape(24)
            boxes3D_corners[ix, :] = camera_to_lidar_cnr(cam_cnr, Tr)      # 8 corners, LiDAR frame
            boxes3D_lidar[ix, :] = lidar_cnr_to_3d(boxes3D_corners[ix, :], lwh[ix, :])
            boxes_bv[ix, :] = lidar_3d_to_bv(boxes3D_lidar[ix, :])
            gt_classes[ix] = cls
            overlaps[ix, cls] = 1.0
        n = ix + 1
        import scipy.sparse
        return {'ry': rys[:n].copy(), 'lwh': lwh[:n].copy(), 'boxes': boxes[:n].copy(), 'boxes_bv': boxes_bv[:n].copy(),
                'boxes_3D_cam': boxes3D[:n].copy(), 'boxes_3D': boxes3D_lidar[:n].copy(),
                'boxes3D_cam_corners': boxes3D_cam_cnr[:n].copy(), 'boxes_corners': boxes3D_corners[:n].copy(),
                'gt_classes': gt_classes[:n].copy(), 'gt_overlaps': scipy.sparse.csr_matrix(overlaps[:n]),
                'xyz': translation[:n].copy(), 'alphas': alphas[:n].copy(), 'flipped': False}

    def _get_obj_level(self, obj):
        height = float(obj[7]) - float(obj[5]) + 1
        trucation = float(obj[1])
        occlusion = float(obj[2])
        if height >= 40 and trucation <= 0.15 and occlusion <= 0:
            return 1
        elif height >= 25 and trucation <= 0.3 and occlusion <= 1:
            return 2
        elif height >= 25 and trucation <= 0.5 and occlusion <= 2:
            return 3
        return 4

    # ------------------------------------------------------------------ frames for the inference loop
    def bev_at(self, i):
        """(H,W,C) float32 BEV map of frame i: the offline .npy if present, else rasterised on the GPU."""
        path = self.lidar_path_at(i)
        if path.endswith('.npy'):
            return np.load(path)
        import torch
        from ..utils.read_lidar import BevRasterizer
        if self._rasterizer is None:
            self._rasterizer = BevRasterizer(**self.raster_args)
        pts = np.fromfile(path, dtype=np.float32).reshape(-1, 4)
        return self._rasterizer(torch.from_numpy(pts).cuda()).cpu().numpy()

    def frame_at(self, i):
        """(raw BGR image HxWx3 float32, BEV HxWxC float32, calib 4x12) -- what test_mv.py:398-403 reads from disk."""
        from ..roi_data_layer.minibatch_mv3d import imread_bgr
        return imread_bgr(self.image_path_at(i)).astype(np.float32), self.bev_at(i), self.calib_at(i)

    # ------------------------------------------------------------------ results (kitti_mv3d.py:321-352,390-395)
    def _write_kitti_results_file(self, all_boxes, all_boxes3D, root=None):
        path = os.path.join(root or ROOT_DIR, 'kitti/results', 'kitti_' + self._subset + '_' + self._image_set + '_'
                            + '-' + time.strftime('%m-%d-%H-%M-%S', time.localtime(time.time())), 'data')
        os.makedirs(path, exist_ok=True)
        for im_ind, index in enumerate(self.image_index):
            filename = os.path.join(path, index + '.txt')
            with open(filename, 'wt') as f:
                for cls_ind, cls in enumerate(self.classes):
                    if cls == '__background__':
                        continue
                    dets = all_boxes[cls_ind][im_ind]
                    if isinstance(dets, list) and dets == []:
                        continue
                    for k in range(dets.shape[0]):
                        alpha = 0
                        f.write('{:s} -1 -1 {:.2f} {:.2f} {:.2f} {:.2f} {:.2f} -1 -1 -1 -1 -1 -1 -1 -1\n'
                                .format(cls.lower(), alpha, dets[k, 0], dets[k, 1], dets[k, 2], dets[k, 3]))
        return path

    def evaluate_detections(self, all_boxes, all_boxes3D, output_dir=None):
        return self._write_kitti_results_file(all_boxes, all_boxes3D, root=output_dir)
