"""get_minibatch (lib/roi_data_layer/minibatch_mv3d.py:17-76): one roidb entry -> the blob dict the MV3D_train
placeholders take.  Same keys, shapes and dtypes.  When the entry's `lidar_bv_path` is a raw Velodyne `.bin` (no offline
raster on disk) the BEV blob is rasterised on the GPU and stays there (a torch CUDA tensor; Network.run takes either)."""
import numpy as np
import numpy.random as npr

from ..fast_rcnn.config import cfg

_RASTER = {}


def imread_bgr(path):
    """cv2.imread(path) of minibatch_mv3d.py:32: HxWx3 uint8 in B,G,R order."""
    from PIL import Image
    return np.ascontiguousarray(np.asarray(Image.open(path).convert('RGB'))[:, :, ::-1])


def load_bev(path, raster_args=None):
    if path.endswith('.npy'):
        return np.load(path)
    import torch
    from ..utils.read_lidar import BevRasterizer
    args = dict(res=0.1, zres=0.3, side_range=(-30., 30.), fwd_range=(0., 60.), height_range=(-2., 0.4))
    args.update(raster_args or {})
    key = tuple(sorted(args.items()))
    if key not in _RASTER:
        _RASTER[key] = BevRasterizer(**args)
    pts = np.fromfile(path, dtype=np.float32).reshape(-1, 4)
    return _RASTER[key](torch.from_numpy(pts).cuda())


def get_minibatch(roidb, num_classes, raster_args=None):
    """Given a roidb (one entry: the reference is single-image), construct a minibatch sampled from it."""
    num_images = len(roidb)
    scales = cfg.TRAIN.get('SCALES', (600,))
    npr.randint(0, high=len(scales), size=num_images)   # minibatch_mv3d.py:22-23: drawn and unused; keeps the RNG stream
    assert cfg.TRAIN.BATCH_SIZE % num_images == 0, \
        'num_images ({}) must divide BATCH_SIZE ({})'.format(num_images, cfg.TRAIN.BATCH_SIZE)
    im_scales = [1]
    im = imread_bgr(roidb[0]['image_path']).astype(np.float32, copy=False)
    lidar_bv_blob = load_bev(roidb[0]['lidar_bv_path'], raster_args)
    im -= cfg.PIXEL_MEANS          # float32 -= float64 (1,1,3): computed in float64, stored float32, as numpy does there
    im_blob = im.reshape((1, im.shape[0], im.shape[1], im.shape[2]))
    lidar_bv_blob = lidar_bv_blob.reshape((1, lidar_bv_blob.shape[0], lidar_bv_blob.shape[1], lidar_bv_blob.shape[2]))
    blobs = {'image_data': im_blob, 'lidar_bv_data': lidar_bv_blob}
    blobs['calib'] = roidb[0]['calib']
    assert len(im_scales) == 1, "Single batch only"
    assert len(roidb) == 1, "Single batch only"
    gt_inds = np.where(roidb[0]['gt_classes'] != 0)[0]
    gt_boxes = np.empty((len(gt_inds), 5), dtype=np.float32)
    gt_boxes[:, 0:4] = roidb[0]['boxes'][gt_inds, :] * im_scales[0]
    gt_boxes[:, 4] = roidb[0]['gt_classes'][gt_inds]
    blobs['gt_boxes'] = gt_boxes
    gt_boxes_bv = np.empty((len(gt_inds), 5), dtype=np.float32)
    gt_boxes_bv[:, 0:4] = roidb[0]['boxes_bv'][gt_inds, :]
    gt_boxes_bv[:, 4] = roidb[0]['gt_classes'][gt_inds]
    blobs['gt_boxes_bv'] = gt_boxes_bv
    gt_boxes_3d = np.empty((len(gt_inds), 7), dtype=np.float32)
    gt_boxes_3d[:, 0:6] = roidb[0]['boxes_3D'][gt_inds, :]
    gt_boxes_3d[:, 6] = roidb[0]['gt_classes'][gt_inds]
    blobs['gt_boxes_3d'] = gt_boxes_3d
    gt_boxes_corners = np.empty((len(gt_inds), 25), dtype=np.float32)
    gt_boxes_corners[:, 0:24] = roidb[0]['boxes_corners'][gt_inds, :]
    gt_boxes_corners[:, 24] = roidb[0]['gt_classes'][gt_inds]
    blobs['gt_boxes_corners'] = gt_boxes_corners
    blobs['im_info'] = np.array([[lidar_bv_blob.shape[1], lidar_bv_blob.shape[2], im_scales[0]]], dtype=np.float32)
    return blobs
