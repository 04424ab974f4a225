"""Generate tests/golden/targets.npz by running the REFERENCE's own anchor_target_layer /
proposal_target_layer_3d (through oracle/ref_shim.py: mechanical py2->py3 patches, arithmetic untouched)
on seeded synthetic inputs.  Dev container only (needs /root/reference):
    python tests/golden/make_golden_targets.py
Inputs are stored next to outputs.  The reference draws its sub-samples from numpy's global RandomState
(seeded like tools/train_net.py:78-80 with cfg.RNG_SEED = 3); the seed used is stored too.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import mv3d_oracle as orc  # noqa: E402  (synthetic input generators only)
from oracle import ref_shim  # noqa: E402


def main():
    R = ref_shim.load()
    R.config.cfg_from_file(R.yml)
    hf, wf = 40, 44
    im_info = np.array([[hf * 8, wf * 8, 1]], dtype=np.float32)
    geom = orc.BevGeometry(0, 32, -17.6, 17.6, 0.1)  # only used to place synthetic cars inside the 320x352 map
    gt_bv, gt_3d, gt_cnr = orc.synth_gt(5, seed=17, geom=orc.REF_GEOMETRY)
    # keep the GT inside the small map: re-centre bv boxes (the layers only ever see bv boxes + 3-D rows)
    gt_bv[:, [0, 2]] = gt_bv[:, [0, 2]] % (wf * 8 - 60) + 5
    gt_bv[:, [1, 3]] = gt_bv[:, [1, 3]] % (hf * 8 - 60) + 5
    gt_bv[:, 2] = np.maximum(gt_bv[:, 2], gt_bv[:, 0] + 14)
    gt_bv[:, 3] = np.maximum(gt_bv[:, 3], gt_bv[:, 1] + 14)
    cls = np.zeros((1, hf, wf, 8), np.float32)
    seed = 3
    np.random.seed(seed)
    labels, targets, anchors, anchors_3d = R.anchor_target_layer_tf.anchor_target_layer(
        cls, gt_bv, gt_3d, im_info, [8, ], [1.0, 1.0])
    prob, deltas = orc.synth_rpn_outputs(hf, wf, seed=23)
    with np.errstate(all="ignore"):
        rois_bv, rois_img, rois_3d = R.proposal_layer_tf.proposal_layer_3d(prob, deltas, im_info, orc.KITTI_CALIB,
                                                                           "TRAIN", [8, ], [1.0, 1.0])
    np.random.seed(seed + 1)
    out = R.proposal_target_layer_tf.proposal_target_layer_3d(rois_bv, rois_3d, gt_bv, gt_3d, gt_cnr, orc.KITTI_CALIB, 2)
    np.savez_compressed(os.path.join(HERE, "targets.npz"), hf=hf, wf=wf, im_info=im_info, gt_bv=gt_bv, gt_3d=gt_3d,
                        gt_cnr=gt_cnr, seed=seed, calib=orc.KITTI_CALIB, at_labels=labels, at_targets=targets,
                        at_anchors=anchors, at_anchors_3d=anchors_3d, rois_bv=rois_bv, rois_3d=rois_3d,
                        pt_rois_bv=out[0], pt_rois_img=out[1], pt_labels=out[2], pt_targets=out[3], pt_rois_3d=out[4])
    print("anchor targets: fg", int((labels == 1).sum()), "bg", int((labels == 0).sum()), "| proposal targets:",
          [o.shape for o in out], "fg", int((out[2] > 0).sum()))


if __name__ == "__main__":
    main()
