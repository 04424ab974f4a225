"""A tiny synthetic KITTI object-detection tree (calib, label_2, image_2, lidar_bv / velodyne) for the feed tests."""
import os

import numpy as np

CALIB = """P0: 7.215377e+02 0.0 6.095593e+02 0.0 0.0 7.215377e+02 1.728540e+02 0.0 0.0 0.0 1.0 0.0
P1: 7.215377e+02 0.0 6.095593e+02 -3.875744e+02 0.0 7.215377e+02 1.728540e+02 0.0 0.0 0.0 1.0 0.0
P2: 7.215377e+02 0.0 6.095593e+02 4.485728e+01 0.0 7.215377e+02 1.728540e+02 2.163791e-01 0.0 0.0 1.0 2.745884e-03
P3: 7.215377e+02 0.0 6.095593e+02 -3.395242e+02 0.0 7.215377e+02 1.728540e+02 2.199936e+00 0.0 0.0 1.0 2.729905e-03
R0_rect: 9.999239e-01 9.837760e-03 -7.445048e-03 -9.869795e-03 9.999421e-01 -4.278459e-03 7.402527e-03 4.351614e-03 9.999631e-01
Tr_velo_to_cam: 7.533745e-03 -9.999714e-01 -6.166020e-04 -4.069766e-03 1.480249e-02 7.280733e-04 -9.998902e-01 -7.631618e-02 9.998621e-01 7.523790e-03 1.480755e-02 -2.717806e-01
Tr_imu_to_velo: 9.999976e-01 7.553071e-04 -2.035826e-03 -8.086759e-01 -7.854027e-04 9.998898e-01 -1.482298e-02 3.195559e-01 2.024406e-03 1.482454e-02 9.998881e-01 -7.997231e-01
"""


def make_tree(root, n_frames=4, seed=0, with_bev=True, with_velodyne=False, hw=(24, 40)):
    rng = np.random.default_rng(seed)
    obj = os.path.join(root, "object", "training")
    for d in ("calib", "label_2", "image_2", "lidar_bv", "velodyne"):
        os.makedirs(os.path.join(obj, d), exist_ok=True)
    os.makedirs(os.path.join(root, "ImageSets"), exist_ok=True)
    from PIL import Image
    index = []
    for i in range(n_frames):
        name = "%06d" % i
        index.append(name)
        # per-frame calib: perturb so that calib_at's position-vs-index rule is observable
        lines = CALIB.strip().split("\n")
        p2 = lines[2].split(" ")
        p2[4] = "%e" % (float(p2[4]) + i)
        lines[2] = " ".join(p2)
        open(os.path.join(obj, "calib", name + ".txt"), "w").write("\n".join(lines) + "\n")
        labs = []
        for k in range(int(rng.integers(1, 5))):
            cls = ["Car", "Car", "Pedestrian", "DontCare", "Van"][int(rng.integers(0, 5))]
            h, w, l = rng.uniform(1.3, 1.8), rng.uniform(1.4, 1.9), rng.uniform(3.2, 4.6)
            tx, ty, tz = rng.uniform(-20, 20), rng.uniform(1.2, 2.0), rng.uniform(6, 55)
            ry = rng.uniform(-np.pi, np.pi)
            x1, y1 = rng.uniform(0, 600), rng.uniform(100, 250)
            labs.append("%s %.2f %d %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f" % (
                cls, rng.uniform(0, 0.4), int(rng.integers(0, 3)), rng.uniform(-3, 3), x1, y1, x1 + rng.uniform(20, 200),
                y1 + rng.uniform(20, 100), h, w, l, tx, ty, tz, ry))
        if i == 1:
            labs = ["Car 0.00 0 1.55 614.24 181.78 727.31 284.77 1.57 1.73 4.15 1.00 1.75 13.22 1.62"] + labs
        open(os.path.join(obj, "label_2", name + ".txt"), "w").write("\n".join(labs) + "\n")
        img = rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
        Image.fromarray(img, "RGB").save(os.path.join(obj, "image_2", name + ".png"))
        if with_bev:
            np.save(os.path.join(obj, "lidar_bv", name + ".npy"), rng.random((33, 41, 9), dtype=np.float32))
        if with_velodyne:
            pts = np.stack([rng.uniform(-5, 65, 3000), rng.uniform(-35, 35, 3000), rng.uniform(-3, 1, 3000),
                            rng.uniform(0, 1, 3000)], 1).astype(np.float32)
            pts.tofile(os.path.join(obj, "velodyne", name + ".bin"))
    # the set file lists a SUBSET in a different order (exercises calib_at's position rule)
    sel = index[1:] if n_frames > 2 else index
    open(os.path.join(root, "ImageSets", "train.txt"), "w").write("\n".join(sel) + "\n")
    return sel
