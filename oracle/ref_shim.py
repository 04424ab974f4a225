"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Runs the *unmodified arithmetic* of the reference (leeyevi/MV3D_TF, mounted read-only
at /root/reference) under Python 3.12 / numpy 2 / Cython 3 so that it can serve as the
pin for `oracle/mv3d_oracle.py` and as the generator of the golden vectors under
`tests/golden/`.

The reference is Python-2 / numpy-1.12 code.  This module copies the handful of hot-path
source files into a scratch directory OUTSIDE the repo (default /tmp/mv3d_ref_shim),
applies purely mechanical py2->py3 / numpy-2 substitutions (listed in `_PY_RULES` /
`_PYX_RULES`; no arithmetic is touched), cythonizes the three .pyx files there, and
imports the result.  Nothing from the reference is ever written into the repository.

It only works where /root/reference exists (the dev container).  On the GPU box it is
absent; tests that need it skip, and everything else relies on the committed goldens.
"""
from __future__ import annotations

import importlib
import os
import re
import shutil
import subprocess
import sys
import types

REF_ROOT = os.environ.get("MV3D_REFERENCE_ROOT", "/root/reference")
SHIM_ROOT = os.environ.get("MV3D_REF_SHIM_DIR", "/tmp/mv3d_ref_shim")

# (relative source under REF_ROOT/lib, relative destination under SHIM_ROOT)
_PY_FILES = [
    ("fast_rcnn/config.py", "fast_rcnn/config.py"),
    ("fast_rcnn/bbox_transform.py", "fast_rcnn/bbox_transform.py"),
    ("fast_rcnn/nms_wrapper.py", "fast_rcnn/nms_wrapper.py"),
    ("rpn_msr/generate_anchors.py", "rpn_msr/generate_anchors.py"),
    ("rpn_msr/proposal_layer_tf.py", "rpn_msr/proposal_layer_tf.py"),
    ("rpn_msr/anchor_target_layer_tf.py", "rpn_msr/anchor_target_layer_tf.py"),
    ("rpn_msr/proposal_target_layer_tf.py", "rpn_msr/proposal_target_layer_tf.py"),
    ("utils/transform.py", "utils/transform.py"),
    ("utils/read_lidar.py", "utils/read_lidar.py"),
]
_PYX_FILES = [
    ("nms/cpu_nms.pyx", "nms/cpu_nms.pyx"),
    ("utils/bbox.pyx", "utils/cython_bbox.pyx"),
    ("utils/nms.pyx", "utils/cython_nms.pyx"),
]

# mechanical source substitutions (regex, replacement)
_PY_RULES = [
    (r"^(\s*)print (?!\()(.*)$", r"\1print(\2)"),          # single-line print statements
    (r"\bxrange\b", "range"),
    (r"\.iteritems\(\)", ".items()"),
    (r"(\w+)\.has_key\((\w+)\)", r"(\2 in \1)"),
    (r"from distutils import spawn", "import shutil as spawn"),
    (r"spawn\.find_executable\(\"nvcc\"\)", "None"),        # force USE_GPU_NMS=False: CPU path is the pin
    (r"yaml\.load\(f\)", "yaml.safe_load(f)"),
    (r"^from generate_anchors import", "from rpn_msr.generate_anchors import"),
    (r"dtype=np\.float\)", "dtype=np.float64)"),
    (r"cfg\.TRAIN\.BATCH_SIZE / num_images", "cfg.TRAIN.BATCH_SIZE // num_images"),
    (r"deltas\.shape\[1\]/24", "deltas.shape[1]//24"),
    (r"corners\.shape\[1\] / 24", "corners.shape[1] // 24"),
    (r"^import matplotlib\.pyplot as plt$", ""),
    (r"^import pdb$", ""),
]
_PYX_RULES = [
    (r"np\.int_t", "np.intp_t"),
    (r"dtype=np\.int\)", "dtype=np.intp)"),
    # `np.float` was Python's float: Cython kept thresh as a Python object and compared
    # PyFloat(ovr) >= thresh, i.e. in double.  C `double` is that comparison; C `float` is NOT.
    (r"np\.float thresh", "double thresh"),
    (r"DTYPE = np\.float$", "DTYPE = np.float64"),
    (r"np\.float_t", "np.float64_t"),
]


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "lib", "rpn_msr"))


def _patch(text: str, rules) -> str:
    out = []
    for line in text.split("\n"):
        for pat, rep in rules:
            line = re.sub(pat, rep, line)
        out.append(line)
    return "\n".join(out)


def _fix_multiline_print(text: str) -> str:
    # anchor_target_layer_tf.py:33-36 has one print statement spanning several lines
    # (inside `if DEBUG:`).  Wrap it so it parses; it never executes (DEBUG = False).
    return re.sub(r"print\(np\.hstack\(\(\)\n(.*?\n)\s*\)\)\n", r"print(np.hstack((\n\1        )))\n", text, flags=re.S)


def build(force: bool = False) -> str:
    """Materialise the patched tree under SHIM_ROOT and compile the Cython modules."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    stamp = os.path.join(SHIM_ROOT, ".built")
    if os.path.exists(stamp) and not force:
        return SHIM_ROOT
    if os.path.isdir(SHIM_ROOT):
        shutil.rmtree(SHIM_ROOT)
    for pkg in ("fast_rcnn", "rpn_msr", "utils", "nms"):
        os.makedirs(os.path.join(SHIM_ROOT, pkg), exist_ok=True)
        open(os.path.join(SHIM_ROOT, pkg, "__init__.py"), "w").close()
    # stub: easydict (15-line dict subclass, not part of the reference)
    with open(os.path.join(SHIM_ROOT, "easydict.py"), "w") as f:
        f.write(
            "class EasyDict(dict):\n"
            "    def __init__(self, d=None, **kw):\n"
            "        super().__init__()\n"
            "        d = dict(d or {}); d.update(kw)\n"
            "        for k, v in d.items():\n"
            "            setattr(self, k, v)\n"
            "    def __setattr__(self, k, v):\n"
            "        if isinstance(v, dict) and not isinstance(v, EasyDict):\n"
            "            v = EasyDict(v)\n"
            "        super().__setitem__(k, v)\n"
            "    __setitem__ = __setattr__\n"
            "    def __getattr__(self, k):\n"
            "        try:\n"
            "            return self[k]\n"
            "        except KeyError:\n"
            "            raise AttributeError(k)\n"
        )
    for src, dst in _PY_FILES:
        text = open(os.path.join(REF_ROOT, "lib", src)).read()
        text = _patch(text, _PY_RULES)
        if src.endswith("anchor_target_layer_tf.py"):
            text = _fix_multiline_print(text)
        if src.endswith("transform.py"):
            text = "from functools import reduce\n" + text
        open(os.path.join(SHIM_ROOT, dst), "w").write(text)
    for src, dst in _PYX_FILES:
        text = open(os.path.join(REF_ROOT, "lib", src)).read()
        open(os.path.join(SHIM_ROOT, dst), "w").write(_patch(text, _PYX_RULES))
    setup_py = os.path.join(SHIM_ROOT, "_setup.py")
    with open(setup_py, "w") as f:
        f.write(
            "from setuptools import setup, Extension\n"
            "from Cython.Build import cythonize\n"
            "import numpy as np\n"
            "exts = [Extension('nms.cpu_nms', ['nms/cpu_nms.pyx'], include_dirs=[np.get_include()]),\n"
            "        Extension('utils.cython_bbox', ['utils/cython_bbox.pyx'], include_dirs=[np.get_include()]),\n"
            "        Extension('utils.cython_nms', ['utils/cython_nms.pyx'], include_dirs=[np.get_include()])]\n"
            "setup(ext_modules=cythonize(exts, language_level=2, quiet=True))\n"
        )
    subprocess.run([sys.executable, "_setup.py", "build_ext", "--inplace", "-q"], cwd=SHIM_ROOT, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    # compile check of every patched python file
    for _, dst in _PY_FILES:
        path = os.path.join(SHIM_ROOT, dst)
        compile(open(path).read(), path, "exec")
    open(stamp, "w").write("ok\n")
    return SHIM_ROOT


def export_binaries(dst: str) -> list:
    """Copy the COMPILED reference Cython modules (binaries only, never sources) to `dst` (oracle/_ref/,
    git-ignored) so that they travel to the GPU box as the reference's own cpu_nms / bbox_overlaps."""
    import glob

    root = build()
    os.makedirs(dst, exist_ok=True)
    out = []
    for so in glob.glob(os.path.join(root, "nms", "*.so")) + glob.glob(os.path.join(root, "utils", "*.so")):
        tgt = os.path.join(dst, os.path.basename(so))
        if not os.path.exists(tgt) or os.path.getmtime(tgt) < os.path.getmtime(so):
            shutil.copy2(so, tgt)
        out.append(tgt)
    return out


_PKGS = ("fast_rcnn", "rpn_msr", "utils", "nms", "easydict")


class _Ref(types.SimpleNamespace):
    pass


def load() -> _Ref:
    """Import the patched reference modules and return them in a namespace.

    The reference uses top-level package names (`utils`, `nms`, ...) that could collide
    with other code, so they are imported with SHIM_ROOT first on sys.path and then
    removed from sys.modules again; the returned namespace keeps them alive.
    """
    root = build()
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _PKGS}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, root)
    try:
        ns = _Ref()
        ns.config = importlib.import_module("fast_rcnn.config")
        ns.cfg = ns.config.cfg
        ns.bbox_transform = importlib.import_module("fast_rcnn.bbox_transform")
        ns.generate_anchors = importlib.import_module("rpn_msr.generate_anchors")
        ns.transform = importlib.import_module("utils.transform")
        ns.read_lidar = importlib.import_module("utils.read_lidar")
        ns.cpu_nms = importlib.import_module("nms.cpu_nms")
        ns.cython_bbox = importlib.import_module("utils.cython_bbox")
        ns.cython_nms = importlib.import_module("utils.cython_nms")
        ns.nms_wrapper = importlib.import_module("fast_rcnn.nms_wrapper")
        ns.proposal_layer_tf = importlib.import_module("rpn_msr.proposal_layer_tf")
        ns.anchor_target_layer_tf = importlib.import_module("rpn_msr.anchor_target_layer_tf")
        ns.proposal_target_layer_tf = importlib.import_module("rpn_msr.proposal_target_layer_tf")
        ns.yml = os.path.join(REF_ROOT, "experiments", "cfgs", "faster_rcnn_end2end.yml")
    finally:
        sys.path.remove(root)
        for k in [k for k in sys.modules if k.split(".")[0] in _PKGS]:
            del sys.modules[k]
        sys.modules.update(saved)
    return ns


if __name__ == "__main__":
    r = load()
    print("reference shim ok:", sorted(vars(r)))


# ---------------------------------------------------------------------------------------------------------------
# KITTI feed: the reference's dataset / data-layer modules, same mechanical treatment
# ---------------------------------------------------------------------------------------------------------------
_FEED_FILES = [
    ("datasets/imdb.py", "datasets/imdb.py"),
    ("datasets/kitti_mv3d.py", "datasets/kitti_mv3d.py"),
    ("roi_data_layer/roidb.py", "roi_data_layer/roidb.py"),
    ("roi_data_layer/minibatch_mv3d.py", "roi_data_layer/minibatch_mv3d.py"),
    ("roi_data_layer/layer.py", "roi_data_layer/layer.py"),
    ("utils/boxes_grid.py", "utils/boxes_grid.py"),
    ("utils/blob.py", "utils/blob.py"),
]
_FEED_RULES = [
    (r"^(\s*)print (?!\()([^#]*?)\s+#(.*)$", r"\1print(\2)  #\3"),   # print statement with a trailing comment
] + _PY_RULES + [
    (r"^import cPickle$", "import pickle as cPickle"),
    (r"^from read_lidar import", "from utils.read_lidar import"),
    (r"if roidb\[i\]\['boxes_corners'\] == \[\]:", "if isinstance(roidb[i]['boxes_corners'], list) and roidb[i]['boxes_corners'] == []:"),
    (r"if dets == \[\]:", "if isinstance(dets, list) and dets == []:"),
]


def build_feed(force: bool = False) -> str:
    root = build(force)
    stamp = os.path.join(root, ".feed_built")
    if os.path.exists(stamp) and not force:
        return root
    for pkg in ("datasets", "roi_data_layer"):
        os.makedirs(os.path.join(root, pkg), exist_ok=True)
    open(os.path.join(root, "roi_data_layer", "__init__.py"), "w").close()
    # datasets/__init__.py of the reference imports every dataset (pascal, coco, matlab ...); the stub keeps the two
    # names kitti_mv3d.py uses: the imdb class and ROOT_DIR (overridable for tests)
    with open(os.path.join(root, "datasets", "__init__.py"), "w") as f:
        f.write("import os\nfrom .imdb import imdb\nROOT_DIR = os.environ.get('MV3D_SHIM_ROOT_DIR', '/tmp/mv3d_ref_shim_root')\n")
    # cv2 is not installed: imread through PIL, in cv2's B,G,R channel order
    with open(os.path.join(root, "cv2.py"), "w") as f:
        f.write("import numpy as np\nfrom PIL import Image\n"
                "def imread(p):\n    return np.ascontiguousarray(np.asarray(Image.open(p).convert('RGB'))[:, :, ::-1])\n")
    for src, dst in _FEED_FILES:
        text = _patch(open(os.path.join(REF_ROOT, "lib", src)).read(), _FEED_RULES)
        open(os.path.join(root, dst), "w").write(text)
        compile(text, os.path.join(root, dst), "exec")
    open(stamp, "w").write("ok\n")
    return root


def load_feed() -> _Ref:
    """kitti_mv3d / prepare_roidb / get_minibatch / RoIDataLayer of the reference, importable under py3."""
    root = build_feed()
    pk = _PKGS + ("datasets", "roi_data_layer", "cv2")
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in pk}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, root)
    try:
        ns = _Ref()
        ns.config = importlib.import_module("fast_rcnn.config")
        ns.cfg = ns.config.cfg
        ns.datasets = importlib.import_module("datasets")
        ns.kitti_mv3d = importlib.import_module("datasets.kitti_mv3d")
        ns.roidb = importlib.import_module("roi_data_layer.roidb")
        ns.minibatch = importlib.import_module("roi_data_layer.minibatch_mv3d")
        ns.layer = importlib.import_module("roi_data_layer.layer")
        ns.transform = importlib.import_module("utils.transform")
        ns.yml = os.path.join(REF_ROOT, "experiments", "cfgs", "faster_rcnn_end2end.yml")
    finally:
        sys.path.remove(root)
        for k in [k for k in sys.modules if k.split(".")[0] in pk]:
            del sys.modules[k]
        sys.modules.update(saved)
    return ns
