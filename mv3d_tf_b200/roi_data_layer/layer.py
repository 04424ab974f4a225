"""RoIDataLayer (lib/roi_data_layer/layer.py:16-70): shuffled walk over the roidb, one minibatch per forward().
Draws its permutations from numpy's global RandomState exactly like the reference (tools/train_net.py:78-80 seeds it)."""
import numpy as np

from ..fast_rcnn.config import cfg
from .minibatch_mv3d import get_minibatch


class RoIDataLayer(object):
    def __init__(self, roidb, num_classes, raster_args=None):
        self._roidb = roidb
        self._num_classes = num_classes
        self._raster_args = raster_args
        self._shuffle_roidb_inds()

    def _shuffle_roidb_inds(self):
        self._perm = np.random.permutation(np.arange(len(self._roidb)))
        self._cur = 0

    def _get_next_minibatch_inds(self):
        # HAS_RPN branch of layer.py:33-38 (the MV3D configuration)
        if self._cur + cfg.TRAIN.IMS_PER_BATCH >= len(self._roidb):
            self._shuffle_roidb_inds()
        db_inds = self._perm[self._cur:self._cur + cfg.TRAIN.IMS_PER_BATCH]
        self._cur += cfg.TRAIN.IMS_PER_BATCH
        return db_inds

    def _get_next_minibatch(self):
        db_inds = self._get_next_minibatch_inds()
        minibatch_db = [self._roidb[i] for i in db_inds]
        return get_minibatch(minibatch_db, self._num_classes, self._raster_args)

    def forward(self):
        return self._get_next_minibatch()

    def __iter__(self):
        while True:
            yield self.forward()
