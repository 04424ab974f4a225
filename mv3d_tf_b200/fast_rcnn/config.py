"""The slice of the reference's global `cfg` that the hot path reads (lib/fast_rcnn/config.py:35-242),
with the live overlay experiments/cfgs/faster_rcnn_end2end.yml applied by `cfg_from_end2end_yml()`.
Same key names; attribute access like easydict."""
from __future__ import annotations

import ast

import numpy as np


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def _defaults() -> AttrDict:
    c = AttrDict()
    c.TRAIN = AttrDict(
        RPN_PRE_NMS_TOP_N=12000, RPN_POST_NMS_TOP_N=2000, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=5,   # config.py:138-147
        RPN_POSITIVE_OVERLAP=0.7, RPN_NEGATIVE_OVERLAP=0.5, RPN_CLOBBER_POSITIVES=False,        # :125-131
        RPN_FG_FRACTION=0.25, RPN_BATCHSIZE=128, RPN_BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0, 1.0, 1.0),  # :132-136,144
        RPN_POSITIVE_WEIGHT=-1.0, BATCH_SIZE=128, FG_FRACTION=0.25, FG_THRESH=0.5,              # :61-70,150
        BG_THRESH_HI=0.5, BG_THRESH_LO=0.1, IMS_PER_BATCH=2, DISPLAY=10, SNAPSHOT_ITERS=5000,
        SCALES=(600,), MAX_SIZE=1000, HAS_RPN=False, USE_FLIPPED=False, SNAPSHOT_PREFIX='VGGnet_fast_rcnn', SNAPSHOT_INFIX='',  # :47-56,84,97-104
        BBOX_NORMALIZE_TARGETS_PRECOMPUTED=False, PROPOSAL_METHOD='selective_search')                 # :92,113
    c.TEST = AttrDict(RPN_PRE_NMS_TOP_N=12000, RPN_POST_NMS_TOP_N=2000, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=5,  # :185-193
                      NMS=0.5, HAS_RPN=False)  # :168,180
    c.PIXEL_MEANS = np.array([[[95.8814, 98.7743, 93.8549]]])  # config.py:211
    c.RNG_SEED = 3
    c.USE_GPU_NMS = True
    c.GPU_ID = 0
    c.EPS = 1e-14
    c.EXP_DIR = 'default'
    c.IS_MULTISCALE = False
    import os.path as osp
    c.ROOT_DIR = osp.abspath(osp.join(osp.dirname(__file__), '..', '..'))   # config.py:218
    return c


cfg = _defaults()


def cfg_from_end2end_yml() -> None:
    """Values of experiments/cfgs/faster_rcnn_end2end.yml:1-20 (the overlay mv3d.sh:35 passes)."""
    cfg.TRAIN.update(RPN_PRE_NMS_TOP_N=12000, RPN_POST_NMS_TOP_N=2000, FG_THRESH=0.7, BG_THRESH_HI=0.5,
                     BG_THRESH_LO=0.0, IMS_PER_BATCH=1, RPN_POSITIVE_OVERLAP=0.7, RPN_BATCHSIZE=128, BATCH_SIZE=128,
                     HAS_RPN=True)
    cfg.TRAIN.update(BBOX_NORMALIZE_TARGETS_PRECOMPUTED=True, PROPOSAL_METHOD='gt')
    cfg.EXP_DIR = 'faster_rcnn_end2end'
    cfg.TEST.update(RPN_PRE_NMS_TOP_N=6000, RPN_POST_NMS_TOP_N=300, NMS=0.1, HAS_RPN=True)


def cfg_from_list(cfg_list) -> None:
    """`--set KEY VALUE ...` overrides (config.py:299-319)."""
    assert len(cfg_list) % 2 == 0
    for k, v in zip(cfg_list[0::2], cfg_list[1::2]):
        keys = k.split(".")
        d = cfg
        for sub in keys[:-1]:
            d = d[sub]
        try:
            value = ast.literal_eval(v)
        except Exception:
            value = v
        assert keys[-1] in d, "unknown config key %s" % k
        d[keys[-1]] = value


def _merge_a_into_b(a, b):
    """config.py:259-287: recursive overlay with key / type checks."""
    for k, v in a.items():
        if k not in b:
            raise KeyError('{} is not a valid config key'.format(k))
        old = b[k]
        if isinstance(old, dict):
            if not isinstance(v, dict):
                raise ValueError('Type mismatch for config key: {}'.format(k))
            _merge_a_into_b(v, old)
            continue
        if isinstance(old, np.ndarray):
            v = np.array(v, dtype=old.dtype)
        elif isinstance(old, tuple) and isinstance(v, list):
            v = tuple(v)
        elif old is not None and type(old) is not type(v) and not (isinstance(old, (int, float)) and isinstance(v, (int, float))):
            raise ValueError('Type mismatch ({} vs. {}) for config key: {}'.format(type(old), type(v), k))
        b[k] = v


def cfg_from_file(filename) -> None:
    """Load a yml overlay such as experiments/cfgs/faster_rcnn_end2end.yml (config.py:289-297)."""
    import yaml
    with open(filename, 'r') as f:
        _merge_a_into_b(yaml.safe_load(f), cfg)


def get_output_dir(imdb, weights_filename=None):
    """<ROOT_DIR>/output/<EXP_DIR>/<imdb.name>[/<weights_filename>], created if missing (config.py:245-257)."""
    import os
    import os.path as osp
    outdir = osp.abspath(osp.join(cfg.ROOT_DIR, 'output', cfg.EXP_DIR, imdb.name))
    if weights_filename is not None:
        outdir = osp.join(outdir, weights_filename)
    if not os.path.exists(outdir):
        os.makedirs(outdir)
    return outdir
