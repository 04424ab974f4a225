"""`RoIDataLayer` -- interface of lib/roi_data_layer/layer.py:16-70 (HAS_RPN configuration, the only one MV3D uses):
`forward()` returns the blob dict of the next training frame.

Order of frames: a random permutation of the roidb, walked `cfg.TRAIN.IMS_PER_BATCH` entries at a time; a fresh
permutation is drawn when fewer than IMS_PER_BATCH + 1 entries remain (so the last entry of an epoch is never served --
the reference's `>=`), and once at construction.  Permutations come from numpy's GLOBAL RandomState, which
tools/train_net.py seeds with cfg.RNG_SEED and which the target layers also draw from: same seed, same frame order and
same sampled anchors as the reference (tests/test_kitti_feed.py compares the order)."""
import numpy as np

from ..fast_rcnn.config import cfg
from .minibatch_mv3d import get_minibatch


class RoIDataLayer(object):
    def __init__(self, roidb, num_classes, raster_args=None):
        self._roidb = roidb
        self._num_classes = num_classes
        self._raster_args = raster_args
        self._new_epoch()

    def _new_epoch(self):
        self._perm = np.random.permutation(np.arange(len(self._roidb)))
        self._cur = 0

    def _next_indices(self):
        step = cfg.TRAIN.IMS_PER_BATCH
        if self._cur + step >= len(self._roidb):
            self._new_epoch()
        first = self._cur
        self._cur = first + step
        return self._perm[first:first + step]

    def forward(self):
        entries = [self._roidb[i] for i in self._next_indices()]
        return get_minibatch(entries, self._num_classes, self._raster_args)

    def __iter__(self):
        while True:
            yield self.forward()
