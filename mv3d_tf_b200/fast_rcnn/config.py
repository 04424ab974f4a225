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
        BG_THRESH_HI=0.5, BG_THRESH_LO=0.1, IMS_PER_BATCH=2, DISPLAY=10, SNAPSHOT_ITERS=5000)
    c.TEST = AttrDict(RPN_PRE_NMS_TOP_N=12000, RPN_POST_NMS_TOP_N=2000, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=5,  # :185-193
                      NMS=0.5)  # :168
    c.PIXEL_MEANS = np.array([[[95.8814, 98.7743, 93.8549]]])  # config.py:211
    c.RNG_SEED = 3
    c.USE_GPU_NMS = True
    c.GPU_ID = 0
    c.EPS = 1e-14
    return c


cfg = _defaults()


def cfg_from_end2end_yml() -> None:
    """Values of experiments/cfgs/faster_rcnn_end2end.yml:1-20 (the overlay mv3d.sh:35 passes)."""
    cfg.TRAIN.update(RPN_PRE_NMS_TOP_N=12000, RPN_POST_NMS_TOP_N=2000, FG_THRESH=0.7, BG_THRESH_HI=0.5,
                     BG_THRESH_LO=0.0, IMS_PER_BATCH=1, RPN_POSITIVE_OVERLAP=0.7, RPN_BATCHSIZE=128, BATCH_SIZE=128)
    cfg.TEST.update(RPN_PRE_NMS_TOP_N=6000, RPN_POST_NMS_TOP_N=300, NMS=0.1)


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
