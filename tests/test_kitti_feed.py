"""the KITTI-MV3D feed (dataset, roidb, minibatch, data layer), the config overlay, the
results writer and the CLI argument surface -- against the reference's own modules run through the shim when
/root/reference exists, and against committed properties otherwise."""
import os

import numpy as np
import pytest

import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kitti_synth import make_tree  # noqa: E402


@pytest.fixture(autouse=True)
def _restore_cfg():
    from mv3d_tf_b200.fast_rcnn import config as c
    yield
    c.cfg.update(c._defaults())


def _ours(root):
    from mv3d_tf_b200.datasets.kitti_mv3d import kitti_mv3d
    return kitti_mv3d('train', kitti_path=root, use_cache=False)


def test_annotation_geometry_properties(tmp_path):
    sel = make_tree(str(tmp_path), n_frames=4)
    d = _ours(str(tmp_path))
    assert d.image_index == sel and d.num_classes == 2 and d.classes == ('__background__', 'Car')
    roidb = d.roidb
    assert len(roidb) == len(sel)
    for e in roidb:
        n = e['boxes'].shape[0]
        assert e['gt_classes'].dtype == np.int32 and (e['gt_classes'] == 1).all()        # only 'Car' rows survive
        for k, shape in (('boxes_bv', (n, 4)), ('boxes_3D', (n, 6)), ('boxes_corners', (n, 24)), ('boxes_3D_cam', (n, 6)),
                         ('boxes3D_cam_corners', (n, 24)), ('lwh', (n, 3)), ('xyz', (n, 3))):
            assert e[k].shape == shape and e[k].dtype == np.float32, k
        assert e['gt_overlaps'].shape == (n, 2) and e['flipped'] is False
        if n:
            # LiDAR box centre = mean of its corners; l,w,h are the label's
            c = e['boxes_corners'].reshape(n, 3, 8)
            assert np.allclose(c.mean(2), e['boxes_3D'][:, :3], atol=1e-5)
            assert np.array_equal(e['boxes_3D'][:, 3:], e['lwh'])
            # BEV box: integral cell indices, x1 <= x2 is NOT guaranteed by the reference (x1 comes from y + w/2)
            assert np.array_equal(e['boxes_bv'], np.round(e['boxes_bv']))
    # the known KITTI label of frame 000001: LiDAR x = camera z, LiDAR y = -camera x, rotated by Tr^-1 only -- the
    # reference drops the translation (zero homogeneous row, transform.py:508-521), so no +0.27 m lever arm
    e = roidb[0]
    assert abs(e['boxes_3D'][0, 0] - 13.22) < 0.1 and abs(e['boxes_3D'][0, 1] - (-1.0)) < 0.15


def test_calib_at_uses_position_not_index(tmp_path):
    make_tree(str(tmp_path), n_frames=4)
    d = _ours(str(tmp_path))
    c = d.calib_at(0)                                   # image_index[0] == '000001' but the file read is 000000.txt
    assert c.shape == (4, 12) and c.dtype == np.float64
    assert c[0, 3] == np.float32(4.485728e+01) and np.all(c[2, 9:] == 0)
    assert d.calib_at(2)[0, 3] == np.float32(4.485728e+01 + 2)


def test_feed_matches_reference(tmp_path, monkeypatch):
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip('reference tree not mounted')
    sel = make_tree(str(tmp_path), n_frames=5, seed=3)
    monkeypatch.setenv('MV3D_SHIM_ROOT_DIR', str(tmp_path / 'refroot'))
    os.makedirs(tmp_path / 'refroot' / 'data' / 'cache', exist_ok=True)
    ref = ref_shim.load_feed()
    ref.config.cfg_from_file(ref.yml)
    ref.cfg.DATA_DIR = str(tmp_path / 'refroot' / 'data')     # the reference caches its roidb pickle under DATA_DIR/cache
    from mv3d_tf_b200.fast_rcnn import config as ours_cfg
    from mv3d_tf_b200.fast_rcnn.train_mv import filter_roidb, get_training_roidb
    from mv3d_tf_b200.roi_data_layer.layer import RoIDataLayer
    from mv3d_tf_b200.roi_data_layer.minibatch_mv3d import get_minibatch
    ours_cfg.cfg_from_file(ref.yml)
    # the yml overlay itself
    for k in ('HAS_RPN', 'IMS_PER_BATCH', 'RPN_POSITIVE_OVERLAP', 'RPN_BATCHSIZE', 'BG_THRESH_LO', 'BG_THRESH_HI', 'FG_THRESH',
              'RPN_PRE_NMS_TOP_N', 'RPN_POST_NMS_TOP_N', 'BATCH_SIZE', 'USE_FLIPPED'):
        assert ours_cfg.cfg.TRAIN[k] == ref.cfg.TRAIN[k], k
    for k in ('RPN_PRE_NMS_TOP_N', 'RPN_POST_NMS_TOP_N', 'NMS', 'HAS_RPN'):
        assert ours_cfg.cfg.TEST[k] == ref.cfg.TEST[k], k
    assert ours_cfg.cfg.EXP_DIR == ref.cfg.EXP_DIR

    rd = ref.kitti_mv3d.kitti_mv3d('train', str(tmp_path))
    od = _ours(str(tmp_path))
    assert od.image_index == rd.image_index and od.name == rd.name and od.classes == rd.classes
    ref.roidb.prepare_roidb(rd)
    ours_roidb = get_training_roidb(od)
    assert len(ours_roidb) == len(rd.roidb) == len(sel)
    for a, b in zip(ours_roidb, rd.roidb):
        assert set(a.keys()) == set(b.keys())
        for k in b:
            if k == 'gt_overlaps':
                assert np.array_equal(a[k].toarray(), b[k].toarray())
            elif isinstance(b[k], np.ndarray):
                assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), k      # bit-exact, incl. dtypes
            else:
                assert a[k] == b[k], k
    assert len(filter_roidb(ours_roidb)) == sum(1 for e in ours_roidb if e['boxes'].shape[0] > 0)   # train_mv.py:348-371
    # minibatch blobs
    for a, b in zip(ours_roidb, rd.roidb):
        mine, theirs = get_minibatch([a], 2), ref.minibatch.get_minibatch([b], 2)
        assert set(mine.keys()) == set(theirs.keys())
        for k in theirs:
            assert mine[k].dtype == theirs[k].dtype and np.array_equal(mine[k], theirs[k]), k
    # data layer: same walk through the roidb from the same seed
    np.random.seed(3)
    lo = RoIDataLayer(ours_roidb, 2)
    seq_o = [lo.forward()['calib'][0, 3] for _ in range(9)]
    np.random.seed(3)
    lr = ref.layer.RoIDataLayer(rd.roidb, 2)
    seq_r = [lr.forward()['calib'][0, 3] for _ in range(9)]
    assert seq_o == seq_r
    # results writer: same file set and contents
    rng = np.random.default_rng(0)
    n = len(sel)
    all_boxes = [[[] for _ in range(n)] for _ in range(2)]
    for i in range(n):
        if i != 1:
            all_boxes[1][i] = rng.uniform(0, 600, (3 + i, 5)).astype(np.float32)
    p_o = od._write_kitti_results_file(all_boxes, None, root=str(tmp_path / 'out_o'))
    ref.datasets.ROOT_DIR = str(tmp_path / 'out_r')
    ref.kitti_mv3d.datasets.ROOT_DIR = str(tmp_path / 'out_r')
    p_r = rd._write_kitti_results_file(all_boxes, None)
    assert sorted(os.listdir(p_o)) == sorted(os.listdir(p_r)) == [s + '.txt' for s in sel]
    for f in os.listdir(p_r):
        assert open(os.path.join(p_o, f)).read() == open(os.path.join(p_r, f)).read()


def test_results_file_format(tmp_path):
    sel = make_tree(str(tmp_path), n_frames=3)
    d = _ours(str(tmp_path))
    all_boxes = [[[] for _ in sel] for _ in range(2)]
    all_boxes[1][0] = np.array([[10.123, 20.5, 30.0, 40.999, 0.9]], np.float32)
    path = d.evaluate_detections(all_boxes, None, output_dir=str(tmp_path / 'o'))
    assert open(os.path.join(path, sel[0] + '.txt')).read() == \
        'car -1 -1 0.00 10.12 20.50 30.00 41.00 -1 -1 -1 -1 -1 -1 -1 -1\n'
    assert open(os.path.join(path, sel[1] + '.txt')).read() == ''


def test_cli_argument_surface():
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for tool, argv, want in (
            ('train_net', ['--device', 'gpu', '--device_id', '0', '--weights', 'w.npy', '--imdb', 'kitti_train', '--iters',
                           '100', '--cfg', 'e.yml', '--network', 'MV3D_train', '--set', 'TRAIN.DISPLAY', '5'],
             dict(device='gpu', device_id=0, pretrained_model='w.npy', imdb_name='kitti_train', max_iters=100,
                  cfg_file='e.yml', network_name='MV3D_train', set_cfgs=['TRAIN.DISPLAY', '5'], randomize=False)),
            ('test_net', ['--device', 'gpu', '--device_id', '1', '--weights', 'm.npy', '--imdb', 'kitti_test', '--cfg',
                          'e.yml', '--network', 'MV3D_test'],
             dict(device='gpu', device_id=1, model='m.npy', imdb_name='kitti_test', cfg_file='e.yml',
                  network_name='MV3D_test', comp_mode=False))):
        spec = importlib.util.spec_from_file_location(tool, os.path.join(root, 'tools', tool + '.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        args = mod.parse_args(argv)
        for k, v in want.items():
            assert getattr(args, k) == v, (tool, k)


def test_cfg_from_list_and_output_dir(tmp_path):
    from mv3d_tf_b200.fast_rcnn import config as c
    old = c.cfg.ROOT_DIR
    try:
        c.cfg.ROOT_DIR = str(tmp_path)
        c.cfg_from_list(['TRAIN.SNAPSHOT_ITERS', '123', 'EXP_DIR', 'exp'])
        assert c.cfg.TRAIN.SNAPSHOT_ITERS == 123

        class I:
            name = 'train'
        out = c.get_output_dir(I(), 'w')
        assert out == os.path.join(str(tmp_path), 'output', 'exp', 'train', 'w') and os.path.isdir(out)
    finally:
        c.cfg.ROOT_DIR = old
        c.cfg.update(c._defaults())


def test_feed_matches_committed_golden(tmp_path, golden_dir):
    """Same comparison against tests/golden/kitti_feed.npz (the reference's outputs on the seed-3 tree, produced by
    tests/golden/make_golden_feed.py) -- runs where the reference tree is not mounted."""
    from mv3d_tf_b200.fast_rcnn import config as ours_cfg
    from mv3d_tf_b200.fast_rcnn.train_mv import get_training_roidb
    from mv3d_tf_b200.roi_data_layer.layer import RoIDataLayer
    from mv3d_tf_b200.roi_data_layer.minibatch_mv3d import get_minibatch

    g = np.load(os.path.join(golden_dir, 'kitti_feed.npz'))
    sel = make_tree(str(tmp_path), n_frames=5, seed=3)
    assert int(g['n']) == len(sel)
    ours_cfg.cfg_from_end2end_yml()
    roidb = get_training_roidb(_ours(str(tmp_path)))
    for i, e in enumerate(roidb):
        for key in [k for k in g.files if k.startswith('roidb%d_' % i)]:
            name = key.split('_', 1)[1]
            mine = e[name].toarray() if name == 'gt_overlaps' else np.asarray(e[name])
            assert mine.dtype == g[key].dtype and np.array_equal(mine, g[key]), key
        blobs = get_minibatch([e], 2)
        for key in [k for k in g.files if k.startswith('blob%d_' % i)]:
            name = key.split('_', 1)[1]
            assert blobs[name].dtype == g[key].dtype and np.array_equal(blobs[name], g[key]), key
    np.random.seed(3)
    layer = RoIDataLayer(roidb, 2)
    assert [layer.forward()['calib'][0, 3] for _ in range(9)] == g['layer_walk'].tolist()
