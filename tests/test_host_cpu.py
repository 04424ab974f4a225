"""CPU-only checks: the C-ABI library loads and exports every symbol include/mv3d_b200.h declares (no compute
calls), the host-side constants (anchor table, projection matrix, raster scalars, config) agree with the oracle,
and the product package never imports the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from mv3d_tf_b200 import build

    return build.build()


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "mv3d_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mv3d_[a-z0-9_]+|_nms)\s*\(", hdr))
    declared -= {"mv3d_status"}
    h = ctypes.CDLL(built_lib)
    missing = [n for n in sorted(declared) if not hasattr(h, n)]
    assert not missing, missing
    from mv3d_tf_b200 import _lib

    assert set(_lib.SIGNATURES) == declared
    h.mv3d_version.restype = ctypes.c_int
    assert h.mv3d_version() >= 100
    h.mv3d_status_string.restype = ctypes.c_char_p
    assert h.mv3d_status_string(-2) == b"workspace too small"


def test_bad_arguments_are_rejected_without_a_gpu(built_lib):
    from mv3d_tf_b200 import _lib

    L = _lib.lib()
    assert L.mv3d_conv_gemm(None, None) == -1
    assert L.mv3d_nms(None, -1, 4, None, 0.7, 1, 0, None, None, None, 0, None) == -1
    assert L.mv3d_bev_raster(None, 10, 4, None, 0, 0, 0, 0, None, None, 0.1, 0, 1, 0, 1, 0, 0, 0, None, 0, None) == -1
    assert L.mv3d_roi_pool_forward(None, 0.125, 5, 0, 0, 0, 7, 7, None, None, None, None) == -1
    assert L.mv3d_proposal_workspace_bytes(None) == 0
    with pytest.raises(_lib.Mv3dError):
        _lib.check(-1, "x")


def test_missing_library_fails_loudly(monkeypatch):
    from mv3d_tf_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmv3d_b200.so")
    with pytest.raises(_lib.Mv3dError):
        _lib.lib()


def test_host_constants_match_oracle(oracle):
    from mv3d_tf_b200.rpn_msr.generate_anchors import all_anchors, generate_anchors_bv
    from mv3d_tf_b200.utils import transform as T
    from mv3d_tf_b200.utils.read_lidar import raster_geometry

    assert np.array_equal(generate_anchors_bv(), oracle.generate_anchors_bv())
    for hf, wf in ((3, 5), (75, 75), (87, 100)):
        a = all_anchors(hf, wf, 8)
        assert np.array_equal(a, oracle.enumerate_anchors(hf, wf, 8))
        assert np.array_equal(T.bv_anchor_to_lidar(a, T.REF_GEOMETRY), oracle.bv_anchor_to_lidar(a, oracle.REF_GEOMETRY))
        assert np.array_equal(T.bv_anchor_to_lidar(a, T.CFG_GEOMETRY), oracle.bv_anchor_to_lidar(a, oracle.CFG_GEOMETRY))
    assert (T.REF_GEOMETRY.xn, T.REF_GEOMETRY.yn, T.CFG_GEOMETRY.xn, T.CFG_GEOMETRY.yn) == (600, 600, 700, 800)
    c = oracle.KITTI_CALIB
    assert np.array_equal(T.projection_matrix(c), oracle.projection_matrix(c[3], c[2], c[0]).astype(np.float32))
    for args in ((0.1, 0.3, (-30., 30.), (0., 60), (-2, 0.4)), (0.1, 0.1, (-40., 40.), (0., 70.), (-2.0, 1.5))):
        g, o = raster_geometry(*args), oracle.raster_params(*args)
        assert (g["H"], g["W"], g["C"], g["nslices"], g["xoff"], g["yoff"]) == \
               (o["H"], o["W"], o["C"], o["nslices"], o["xoff"], o["yoff"])
        assert np.array_equal(g["lo"], o["lo"]) and np.array_equal(g["hi"], o["hi"])
    assert raster_geometry(0.1, 0.1, (-40., 40.), (0., 70.), (-2.0, 1.5))["H"] == 701


def test_config_overlay_and_set():
    from mv3d_tf_b200.fast_rcnn import config as C

    C.cfg_from_end2end_yml()
    assert (C.cfg.TEST.RPN_PRE_NMS_TOP_N, C.cfg.TEST.RPN_POST_NMS_TOP_N) == (6000, 300)
    assert (C.cfg.TRAIN.RPN_PRE_NMS_TOP_N, C.cfg.TRAIN.RPN_POST_NMS_TOP_N, C.cfg.TRAIN.FG_THRESH) == (12000, 2000, 0.7)
    C.cfg_from_list(["TEST.RPN_POST_NMS_TOP_N", "100"])
    assert C.cfg.TEST.RPN_POST_NMS_TOP_N == 100
    C.cfg_from_end2end_yml()
    with pytest.raises(AssertionError):
        C.cfg_from_list(["TEST.NOPE", "1"])


def test_graph_builder_wiring_without_gpu():
    """The reference's layer names / shapes exist; building the graph needs no device."""
    from mv3d_tf_b200.networks.factory import get_network

    net = get_network("MV3D_test", bv_channels=36, device="cpu")
    for name in ("conv5_3", "conv5_3_2", "rpn_cls_prob_reshape", "rpn_bbox_pred", "rois", "roi_data_bv", "roi_data_img",
                 "pool_5", "pool_5_2", "fc7_1", "fc7_2", "cls_prob", "bbox_pred"):
        assert net.get_output(name) is not None
    s = net.param_specs
    assert s["conv1_1"]["shape"] == (3, 3, 36, 64) and s["conv1_1_2"]["shape"] == (3, 3, 3, 64)
    assert s["fc6_1"]["shape"] == (25088, 2048) and s["bbox_pred"]["shape"] == (4096, 48)
    assert s["bbox_pred"]["stddev"] == 0.001 and s["rpn_cls_score"]["shape"] == (1, 1, 512, 8)
    assert isinstance(net.get_output("rois"), tuple) and len(net.get_output("rois")) == 4
    with pytest.raises(KeyError):
        net.feed("nope")
    n_params = sum(int(np.prod(v["shape"])) + v["shape"][-1] for v in s.values())
    assert 140e6 < n_params < 150e6  # ~143 M


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "mv3d_tf_b200")
    bad = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b|from\s+\.\.?oracle|oracle/", txt, flags=re.M):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_mixed_mode_format_plan_is_static_graph_analysis():
    """Which conv outputs travel as f16e5 is decided from the graph alone (no device needed): every PAD tensor whose
    readers are 3x3 convs (directly or through 2x2 pools) -- the trunks; bf16 hi/lo wherever a 1x1 conv reads it
    (rpn_conv/3x3) or nobody reads the PAD rendering (conv5_3_2 feeds only the ROI pool); never while training."""
    from mv3d_tf_b200 import kernels as K
    from mv3d_tf_b200.networks.factory import get_network

    net = get_network("MV3D_test", bv_channels=36, precise=True, mixed=True, fv=True)
    fm = {n.name: net._pad_out_fmt(n) for n in net._program if n.kind == "conv"}
    for s in ("", "_2", "_3"):
        for name in ("conv1_1", "conv1_2", "conv2_2", "conv3_3", "conv4_3", "conv5_2"):
            assert fm[name + s] == K.FMT_F16E5, name + s
    assert fm["conv5_3"] == K.FMT_F16E5            # read by rpn_conv/3x3 (a 3x3 conv)
    assert fm["conv5_3_2"] == fm["conv5_3_3"] == K.FMT_BF16X2
    assert fm["rpn_conv/3x3"] == fm["rpn_cls_score"] == fm["rpn_bbox_pred"] == K.FMT_BF16X2
    net.training = True
    assert all(net._pad_out_fmt(n) == K.FMT_BF16X2 for n in net._program if n.kind == "conv")
    plain = get_network("MV3D_test", bv_channels=36, precise=True)
    assert all(plain._pad_out_fmt(n) == K.FMT_BF16X2 for n in plain._program if n.kind == "conv")
    assert not get_network("MV3D_test", precise=False, mixed=True).mixed     # mixed is a parity mode: needs precise


def test_gemm_dispatch_mirror():
    """kernels.gemm_kernel_name mirrors the C dispatch (bench.py attributes GEMM time by it)."""
    from mv3d_tf_b200 import kernels as K

    old = K.PAIR_MODE
    try:
        K.PAIR_MODE = True
        assert K.gemm_kernel_name(9, 512, 512, 2) == "conv3x3_pair_kernel<256,2>"
        assert K.gemm_kernel_name(9, 64, 64, 3) == "conv3x3_pair_kernel<64,3>"
        assert K.gemm_kernel_name(9, 64, 128, 2) == "conv3x3_pair_kernel<128,2>"
        assert K.gemm_kernel_name(9, 64, 96, 3) == "conv3x3_reuse_kernel<128,3>"      # width does not tile a pair
        assert K.gemm_kernel_name(1, 25088, 2048, 3, split_k=3, m=300) == "fc_swapped_pair_kernel"
        assert K.gemm_kernel_name(1, 25088, 2048, 3, split_k=3, m=2000) == "conv_gemm_kernel<128,64,3>"
        assert K.gemm_kernel_name(1, 32, 64, 3) == "conv_gemm_kernel<64,32,3>"
        K.PAIR_MODE = False
        assert K.gemm_kernel_name(9, 512, 512, 3) == "conv3x3_reuse_kernel<128,3>"
        assert K.gemm_kernel_name(9, 512, 512, 2) == "conv3x3_reuse_kernel<128,2>"
    finally:
        K.PAIR_MODE = old


def test_ctypes_signatures_match_header_prototypes():
    """Every ctypes argtypes list has the arity and the pointer / integer / float kinds of its prototype in
    include/mv3d_b200.h (a mismatch would otherwise only surface as garbage arguments on the GPU box), and the two
    descriptor structs have the header's field order."""
    import ctypes as C
    from mv3d_tf_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "mv3d_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = dict(re.findall(r"\b(mv3d_[a-z0-9_]+|_nms)\s*\(([^;{}]*)\)\s*;", hdr))

    def kind_of_decl(p):
        p = p.strip()
        if "*" in p:
            return "ptr"
        t = p.split()
        base = " ".join(t[:-1]) if len(t) > 1 else t[0]
        if "float" in base or "double" in base:
            return "float"
        return "int"

    def kind_of_ctype(t):
        if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return "ptr"
        if t in (C.c_float, C.c_double):
            return "float"
        return "int"

    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert name in protos, name
        params = [p for p in (x.strip() for x in protos[name].replace("\n", " ").split(",")) if p and p != "void"]
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
        assert [kind_of_decl(p) for p in params] == [kind_of_ctype(t) for t in argtypes], name
    for struct, cls in (("mv3d_gemm_desc", _lib.GemmDesc), ("mv3d_wgrad_desc", _lib.WgradDesc)):
        end = re.search(r"\}\s*%s;" % struct, hdr).start()
        body = hdr[hdr.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        fields = []
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            for part in stmt.split(","):
                fields.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
        assert fields == [f[0] for f in cls._fields_], struct
