"""`prepare_roidb(imdb)` -- interface of lib/roi_data_layer/roidb.py:16-58.  Adds to every roidb entry what the trainer
reads later: where the frame's files are (`image_path`, `lidar_bv_path`), its `calib`, and per annotation the best
class (`max_classes`) and its overlap (`max_overlaps`) from the sparse `gt_overlaps` matrix (`filter_roidb` in
fast_rcnn/train_mv.py keeps entries that have a foreground or background roi by these).  Entries whose corner boxes
are an empty *list* (the imdb's marker for a frame with no usable annotation) are reported and left untouched, as the
reference does."""
import numpy as np


def _class_bookkeeping(gt_overlaps):
    dense = gt_overlaps.toarray()
    best, cls = dense.max(axis=1), dense.argmax(axis=1)
    # a roi without overlap is background (class 0); one with overlap has a foreground class
    if cls[best == 0].any() or not cls[best > 0].all():
        raise AssertionError('gt_overlaps inconsistent with the class labels')
    return cls, best


def prepare_roidb(imdb):
    for i, entry in enumerate(imdb.roidb[:len(imdb.image_index)]):
        corners = entry['boxes_corners']
        if isinstance(corners, list) and len(corners) == 0:
            print('boxes_corners not correct', imdb.image_path_at(i))
            continue
        entry.update(image_path=imdb.image_path_at(i), lidar_bv_path=imdb.lidar_path_at(i), calib=imdb.calib_at(i))
        entry['max_classes'], entry['max_overlaps'] = _class_bookkeeping(entry['gt_overlaps'])
