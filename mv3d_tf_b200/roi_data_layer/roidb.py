"""prepare_roidb (lib/roi_data_layer/roidb.py:16-58): enrich the imdb's roidb with paths, calib and the max-overlap
class bookkeeping the trainer's filter reads."""
import numpy as np


def prepare_roidb(imdb):
    roidb = imdb.roidb
    for i in range(len(imdb.image_index)):
        if len(roidb[i]['boxes_corners']) == 0 and isinstance(roidb[i]['boxes_corners'], list):
            print('boxes_corners not correct', imdb.image_path_at(i))
            continue
        roidb[i]['image_path'] = imdb.image_path_at(i)
        roidb[i]['lidar_bv_path'] = imdb.lidar_path_at(i)
        roidb[i]['calib'] = imdb.calib_at(i)
        gt_overlaps = roidb[i]['gt_overlaps'].toarray()
        max_overlaps = gt_overlaps.max(axis=1)
        max_classes = gt_overlaps.argmax(axis=1)
        roidb[i]['max_classes'] = max_classes
        roidb[i]['max_overlaps'] = max_overlaps
        zero_inds = np.where(max_overlaps == 0)[0]
        assert all(max_classes[zero_inds] == 0)
        nonzero_inds = np.where(max_overlaps > 0)[0]
        assert all(max_classes[nonzero_inds] != 0)
