"""Blob assembly for one training frame.  Interface of lib/roi_data_layer/minibatch_mv3d.py:17-76 (`get_minibatch(roidb,
num_classes)` -> dict), written from the blob specification below rather than from the reference text:

    image_data        (1, H, W, 3)  float32   BGR image minus cfg.PIXEL_MEANS
    lidar_bv_data     (1, Hb, Wb, C) float32  BEV map: the offline `.npy` raster, or -- new here -- rasterised on the GPU
                                              from the raw Velodyne `.bin` (then a CUDA tensor; Network.run takes either)
    calib             (4, 12)                 P2 / P3 / R0 / Tr rows as kitti_mv3d.calib_at returns them
    gt_boxes          (G, 4+1)  float32       image boxes | class      } rows = annotations whose class != 0,
    gt_boxes_bv       (G, 4+1)  float32       BEV boxes | class        } in roidb order
    gt_boxes_3d       (G, 6+1)  float32       x,y,z,l,w,h | class      }
    gt_boxes_corners  (G, 24+1) float32       8 corners (x0..7,y0..7,z0..7) | class
    im_info           (1, 3)    float32       [Hb, Wb, scale = 1]

The reference draws (and discards) one `npr.randint` per call for its unused multi-scale option (:22-23); the draw is
kept so that numpy's global random stream -- which the target layers sample from afterwards -- stays in step with it.
Bit-identical blobs vs the reference run through the shim: tests/test_kitti_feed.py."""
import numpy as np
import numpy.random as npr

from ..fast_rcnn.config import cfg

# blob name -> (roidb field, columns before the class label)
GT_BLOBS = {'gt_boxes': ('boxes', 4), 'gt_boxes_bv': ('boxes_bv', 4), 'gt_boxes_3d': ('boxes_3D', 6),
            'gt_boxes_corners': ('boxes_corners', 24)}
IMAGE_SCALE = 1            # the MV3D feed never rescales the image
REF_RASTER = dict(res=0.1, zres=0.3, side_range=(-30., 30.), fwd_range=(0., 60.), height_range=(-2., 0.4))
_rasterizers = {}


def imread_bgr(path):
    """What cv2.imread returns for a colour image: (H, W, 3) uint8, channels in B,G,R order."""
    from PIL import Image
    rgb = np.asarray(Image.open(path).convert('RGB'))
    return np.ascontiguousarray(rgb[:, :, ::-1])


def load_bev(path, raster_args=None):
    """BEV map of one frame: `.npy` = the reference's offline raster; anything else is a Velodyne float32 (n,4) file that
    is rasterised on the GPU (one cached BevRasterizer per grid configuration)."""
    if path.endswith('.npy'):
        return np.load(path)
    import torch
    from ..utils.read_lidar import BevRasterizer
    grid = dict(REF_RASTER, **(raster_args or {}))
    key = tuple(sorted(grid.items()))
    raster = _rasterizers.get(key)
    if raster is None:
        raster = _rasterizers[key] = BevRasterizer(**grid)
    cloud = np.fromfile(path, dtype=np.float32).reshape(-1, 4)
    return raster(torch.from_numpy(cloud).cuda())


def _labelled_rows(entry, field, width, rows, classes, scale=None):
    out = np.empty((rows.size, width + 1), dtype=np.float32)
    vals = entry[field][rows, :]
    out[:, :width] = vals if scale is None else vals * scale
    out[:, width] = classes
    return out


def get_minibatch(roidb, num_classes, raster_args=None):
    n = len(roidb)
    npr.randint(0, high=len(cfg.TRAIN.get('SCALES', (600,))), size=n)     # see the module docstring
    if cfg.TRAIN.BATCH_SIZE % n != 0:
        raise AssertionError('num_images ({}) must divide BATCH_SIZE ({})'.format(n, cfg.TRAIN.BATCH_SIZE))
    if n != 1:
        raise AssertionError('Single batch only')
    entry = roidb[0]
    image = imread_bgr(entry['image_path']).astype(np.float32, copy=False)
    image -= cfg.PIXEL_MEANS       # float32 array, float64 means: numpy computes in float64 and stores float32
    bev = load_bev(entry['lidar_bv_path'], raster_args)
    blobs = {'image_data': image[None], 'lidar_bv_data': bev[None], 'calib': entry['calib']}
    fg = np.flatnonzero(entry['gt_classes'] != 0)
    classes = entry['gt_classes'][fg]
    for name, (field, width) in GT_BLOBS.items():
        blobs[name] = _labelled_rows(entry, field, width, fg, classes, IMAGE_SCALE if name == 'gt_boxes' else None)
    blobs['im_info'] = np.array([[bev.shape[0], bev.shape[1], IMAGE_SCALE]], dtype=np.float32)
    return blobs
