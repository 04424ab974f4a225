"""BEV anchors (host constants).  Mirrors lib/rpn_msr/generate_anchors.py:37-51 and the shift
enumeration of proposal_layer_tf.py:79-95 / anchor_target_layer_tf.py:76-89."""
import numpy as np


def generate_anchors_bv(base_size=((3.9, 1.6), (1.0, 0.6)), res=0.1):
    boxes = []
    for length, width in base_size:
        nl, nw = int(length / res), int(width / res)
        boxes.append((-(nl // 2), -(nw // 2), nl - nl // 2, nw - nw // 2))
    a = np.array(boxes, dtype=np.int64)
    return np.concatenate([a, a[:, [1, 0, 3, 2]]], axis=0)


def all_anchors(height, width, feat_stride=8):
    """(height*width*A, 4) int64 in (h, w, a) order."""
    base = generate_anchors_bv()
    ys, xs = np.meshgrid(np.arange(height, dtype=np.int64) * feat_stride,
                         np.arange(width, dtype=np.int64) * feat_stride, indexing="ij")
    shift = np.stack([xs, ys, xs, ys], axis=-1).reshape(-1, 1, 4)
    return (shift + base.reshape(1, -1, 4)).reshape(-1, 4)
