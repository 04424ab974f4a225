"""Generate tests/golden/kitti_feed.npz by running the REFERENCE's dataset / data-layer modules (lib/datasets/kitti_mv3d.py,
lib/roi_data_layer/{roidb,minibatch_mv3d,layer}.py through oracle/ref_shim.load_feed: mechanical py2->py3 patches,
arithmetic untouched) over the synthetic KITTI tree of tests/kitti_synth.py (seed 3, 5 frames).

Run in the dev container only (needs /root/reference):   python tests/golden/make_golden_feed.py
The tree itself is regenerated from its seed by the test; only the reference's OUTPUTS are stored.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from kitti_synth import make_tree  # noqa: E402
from oracle import ref_shim  # noqa: E402

ARRAY_KEYS = ('ry', 'lwh', 'boxes', 'boxes_bv', 'boxes_3D_cam', 'boxes_3D', 'boxes3D_cam_corners', 'boxes_corners',
              'gt_classes', 'xyz', 'alphas', 'calib', 'max_classes', 'max_overlaps')
BLOB_KEYS = ('image_data', 'lidar_bv_data', 'calib', 'gt_boxes', 'gt_boxes_bv', 'gt_boxes_3d', 'gt_boxes_corners', 'im_info')


def main():
    tmp = tempfile.mkdtemp(prefix='mv3d_feed_golden_')
    sel = make_tree(tmp, n_frames=5, seed=3)
    os.environ['MV3D_SHIM_ROOT_DIR'] = os.path.join(tmp, 'refroot')
    ref = ref_shim.load_feed()
    ref.config.cfg_from_file(ref.yml)
    ref.cfg.DATA_DIR = os.path.join(tmp, 'refroot', 'data')
    d = ref.kitti_mv3d.kitti_mv3d('train', tmp)
    ref.roidb.prepare_roidb(d)
    out = {'n': np.int64(len(sel))}
    for i, e in enumerate(d.roidb):
        for k in ARRAY_KEYS:
            out['roidb%d_%s' % (i, k)] = np.asarray(e[k])
        out['roidb%d_gt_overlaps' % i] = e['gt_overlaps'].toarray()
        blobs = ref.minibatch.get_minibatch([e], 2)
        for k in BLOB_KEYS:
            out['blob%d_%s' % (i, k)] = np.asarray(blobs[k])
    np.random.seed(3)
    layer = ref.layer.RoIDataLayer(d.roidb, 2)
    out['layer_walk'] = np.array([layer.forward()['calib'][0, 3] for _ in range(9)])
    np.savez_compressed(os.path.join(HERE, 'kitti_feed.npz'), **out)
    print('kitti_feed.npz:', len(out), 'arrays,', os.path.getsize(os.path.join(HERE, 'kitti_feed.npz')), 'bytes')


if __name__ == '__main__':
    main()
