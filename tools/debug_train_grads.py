"""Debug: per-layer gradient error of the GPU train step vs the oracle in fp32 and fp64 (run under gpurun)."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import build as ob
ob.build()
from oracle import mv3d_oracle as oracle, net_oracle
import test_gpu_train as T
from mv3d_tf_b200.fast_rcnn.config import cfg, cfg_from_end2end_yml
from mv3d_tf_b200.fast_rcnn.train_mv import SolverWrapper, _node

cfg_from_end2end_yml(); cfg.USE_GPU_NMS = False
B = 1
net, frames, blobs, ogeom = T._make_problem(oracle, B)
sw = SolverWrapper(network=net, keep_prob=1.0, lr=1e-3)
params0 = sw.export_params()
np.random.seed(3)
loss = sw.train_step(blobs, keep_prob=1.0, apply_update=False)
torch.cuda.synchronize()
vals = net.last_vals
grads = sw.export_grads()
rd = vals[_node(net, "roi_data_3d")].extra
ad = vals[_node(net, "rpn_data")].extra
rois_bv = rd["bv"].cpu().numpy().copy(); rois_img = rd["img"].cpu().numpy().copy()
teacher = dict(rpn_data=(ad["labels"][0].cpu().numpy(), ad["targets"][0].cpu().numpy()),
               roi_data=(rois_bv, rois_img, rd["labels"].cpu().numpy(), rd["targets"].cpu().numpy()))
f = frames[0]
res = {}
for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
    res[name] = net_oracle.train_forward_backward(f["bv"][None], f["img"][None], blobs["im_info"], blobs["calib"], *f["gt"],
                                                  params0, geom=ogeom, teacher=teacher, dtype=dt)
print("loss ours", loss.cpu().numpy(), "f32", res["f32"][0], "f64", res["f64"][0])
print("%-22s %10s %10s %10s" % ("param", "ours-f64", "f32-f64", "ours-f32"))
for k in res["f64"][1]:
    for kk in ("weights", "biases"):
        r64 = res["f64"][1][k][kk]; r32 = res["f32"][1][k][kk]; o = grads[k][kk]
        m = max(np.abs(r64).max(), 1e-30)
        print("%-22s %10.2e %10.2e %10.2e  max|g|=%.2e" % (k + "/" + kk[0], np.abs(o - r64).max() / m, np.abs(r32 - r64).max() / m,
                                             np.abs(o - r32).max() / m, m))

# ---- hypothesis: the residual is ReLU-gate flips at units whose pre-activation is within the forward error of zero
keep = {}
net_oracle.trunk(f["bv"][None], params0, "", torch.float64, keep)
print("%-10s %8s %10s %12s %12s" % ("layer", "flips", "units", "fwd maxerr", "grad L2 rel"))
for name, act in keep.items():
    ours = sw.net.last_vals[_node(net, name)].pad
    from mv3d_tf_b200 import kernels as K
    o = K.unpad_nhwc(ours).cpu().double()
    flips = int(((o > 0) != (act > 0)).sum())
    fe = float((o - act).abs().max() / act.abs().max())
    r64 = res["f64"][1][name]["weights"]; og = grads[name]["weights"]
    l2 = float(np.linalg.norm(og - r64) / np.linalg.norm(r64))
    print("%-10s %8d %10d %12.2e %12.2e" % (name, flips, act.numel(), fe, l2))
