"""SURVEY 8f rows 2-4 on the GPU: the CLI entry points (tools/train_net.py, tools/test_net.py) over a synthetic KITTI
tree whose BEV maps are rasterised ONLINE from raw Velodyne clouds, the snapshot written in the reference's `.npy`
layer-dict format and loaded back, and KITTI-format result files."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kitti_synth import make_tree  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tool(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, 'tools', name + '.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_online_raster_feed_equals_oracle(tmp_path, oracle):
    from mv3d_tf_b200.datasets.kitti_mv3d import kitti_mv3d
    make_tree(str(tmp_path), n_frames=3, with_bev=False, with_velodyne=True, hw=(96, 320))
    d = kitti_mv3d('train', kitti_path=str(tmp_path), use_cache=False)
    im, bv, calib = d.frame_at(0)
    assert im.shape == (96, 320, 3) and im.dtype == np.float32 and calib.shape == (4, 12)
    pts = np.fromfile(d.velodyne_path_from_index(d.image_index[0]), dtype=np.float32).reshape(-1, 4)
    assert bv.shape == (601, 601, 9)                                   # the reference's offline raster shape
    assert np.array_equal(bv, oracle.point_cloud_2_top(pts, **d.raster_args))


def test_train_then_test_through_the_cli(tmp_path):
    from mv3d_tf_b200.fast_rcnn import config as c
    kitti = tmp_path / 'KITTI'
    sel = make_tree(str(kitti), n_frames=3, with_bev=False, with_velodyne=True, hw=(96, 320), seed=5)
    # every training frame needs a car the RPN can anchor on: the fixed KITTI label of kitti_synth
    for s in sel:
        open(kitti / 'object' / 'training' / 'label_2' / (s + '.txt'), 'w').write(
            "Car 0.00 0 1.55 114.24 41.78 227.31 84.77 1.57 1.73 4.15 1.00 1.75 13.22 1.62\n"
            "Car 0.00 0 -1.2 14.24 31.78 127.31 74.77 1.50 1.60 3.90 -6.00 1.70 25.00 -1.2\n")
    old_root = c.cfg.ROOT_DIR
    try:
        c.cfg.ROOT_DIR = str(tmp_path)
        yml = tmp_path / 'e2e.yml'
        yml.write_text("EXP_DIR: faster_rcnn_end2end\nTRAIN:\n  HAS_RPN: True\n  IMS_PER_BATCH: 1\n  RPN_POSITIVE_OVERLAP: 0.7\n"
                       "  RPN_BATCHSIZE: 128\n  BG_THRESH_LO: 0.0\n  BG_THRESH_HI : 0.5\n  FG_THRESH : 0.7\n"
                       "  RPN_PRE_NMS_TOP_N : 2000\n  RPN_POST_NMS_TOP_N : 300\n  SNAPSHOT_ITERS: 2\n"
                       "TEST:\n  RPN_PRE_NMS_TOP_N : 1000\n  RPN_POST_NMS_TOP_N : 100\n  HAS_RPN: True\n  NMS : 0.1\n")
        sw = _tool('train_net').main(['--device', 'gpu', '--device_id', '0', '--imdb', 'kitti_train', '--iters', '2', '--cfg',
                                      str(yml), '--network', 'MV3D_train', '--kitti', str(kitti)])
        snap = os.path.join(str(tmp_path), 'output', 'faster_rcnn_end2end', 'train', 'VGGnet_fast_rcnn_iter_2.npy')
        assert os.path.exists(snap)
        saved = np.load(snap, allow_pickle=True).item()
        live = sw.export_params()
        assert set(saved) == set(live) and saved['conv1_1']['weights'].shape == (3, 3, 9, 64)
        assert saved['fc6_1']['weights'].shape == (25088, 2048) and saved['bbox_pred']['weights'].shape == (4096, 48)
        for k in ('conv1_1', 'conv5_3_2', 'rpn_conv/3x3', 'fc7_2', 'cls_score'):
            assert np.array_equal(saved[k]['weights'], live[k]['weights']) and np.array_equal(saved[k]['biases'], live[k]['biases'])
        del sw
        torch.cuda.empty_cache()
        out = _tool('test_net').main(['--device', 'cpu', '--weights', snap, '--imdb', 'kitti_train', '--cfg', str(yml),
                                      '--network', 'MV3D_test', '--kitti', str(kitti)])
        files = sorted(os.listdir(out))
        assert files == [s + '.txt' for s in sel]
        for f in files:
            for line in open(os.path.join(out, f)):
                tok = line.split(' ')
                assert tok[0] == 'car' and len(tok) == 16 and tok[1] == '-1' and float(tok[4]) == float(tok[4])
    finally:
        c.cfg.ROOT_DIR = old_root
        c.cfg.update(c._defaults())
