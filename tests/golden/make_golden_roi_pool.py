"""Generate tests/golden/roi_pool.npz with the REFERENCE's own RoiPool / RoiPoolGrad CPU kernels
(lib/roi_pooling_layer/roi_pooling_op.cc compiled unmodified by oracle/build_ref_roi_pool.py).
Dev container only (needs /root/reference):   python tests/golden/make_golden_roi_pool.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_roi_pool  # noqa: E402


def main():
    rng = np.random.default_rng(21)
    B, H, W, C, R = 2, 23, 37, 16, 48
    data = rng.normal(size=(B, H, W, C)).astype(np.float32)
    data[0, 3, :, :] = 0.75                       # ties: the first maximum in (h, w) order wins
    x1, y1 = rng.integers(-60, W * 8, R), rng.integers(-60, H * 8, R)
    rois = np.stack((rng.integers(0, B, R), x1, y1, x1 + rng.integers(0, 300, R), y1 + rng.integers(0, 200, R)), 1)
    rois = rois.astype(np.float32)
    rois[0] = [0, -500, -500, -300, -300]         # entirely outside: zeros, argmax -1
    rois[1] = [0, 40, 40, 8, 8]                   # malformed (end < start): forced 1x1
    rois[2, 1:] += 0.5                            # x.5 * 0.125 products: round half away from zero
    rois[3] = [1, 4, 4, 4, 4]                     # single cell
    rois[4] = [1, -4, -4, 3, 3]                   # negative start: round(-0.5) = -1
    top, arg = ref_roi_pool.roi_pool_forward(data, rois)
    grad = rng.normal(size=top.shape).astype(np.float32)
    dbottom = ref_roi_pool.roi_pool_backward(data, rois, arg, grad)
    np.savez_compressed(os.path.join(HERE, "roi_pool.npz"), data=data, rois=rois, top=top, argmax=arg, grad=grad,
                        dbottom=dbottom)
    print("roi_pool", top.shape, int((arg < 0).sum()), float(np.abs(dbottom).sum()))


if __name__ == "__main__":
    main()
