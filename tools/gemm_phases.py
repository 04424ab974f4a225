#!/usr/bin/env python
"""Where one 3x3 pair-kernel launch spends its time: clock64 stamps of CTA pair 0 (mv3d_gemm_set_stamps) for the layer
shapes of the BASELINE trunks, f16e5 (2-pass) and bf16x3 (3-pass) operands.  Measurement tool, not a product path.
    python tools/gemm_phases.py [out.md]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mv3d_tf_b200 import kernels as K  # noqa: E402
from mv3d_tf_b200._lib import check, lib, ptr  # noqa: E402

PHASES = ["set-up", "first operands", "MMA issue", "MMA drain", "epilogue", "other warps", "teardown"]


def main():
    shapes = [("conv1_2 BEV", 701, 801, 64, 64), ("conv2_2 BEV", 350, 400, 128, 128), ("conv3_1 BEV", 175, 200, 128, 256),
              ("conv3_3 BEV", 175, 200, 256, 256), ("conv4_1 BEV", 87, 100, 256, 512), ("conv5_x BEV", 87, 100, 512, 512),
              ("conv5_x RGB", 46, 155, 512, 512)]
    stamps = torch.zeros(8, dtype=torch.int64, device="cuda")
    lines = ["# conv3x3_pair_kernel phase breakdown (CTA pair 0, clock64 cycles -> us at the SM clock read from nvidia-smi max)", "",
             "| layer | fmt | total us (events) | " + " | ".join(PHASES) + " | sum us |", "|---|---|---|" + "---|" * (len(PHASES) + 1)]
    mhz = 1965.0
    for name, H, W, cin, cout in shapes:
        for fmt in (K.FMT_F16E5, K.FMT_BF16X2):
            x = torch.randn((1, H, W, cin), device="cuda") * 0.5
            a = K.pad_nhwc(x, precise=True, fmt=fmt)
            w = K.pack_weights(torch.randn((3, 3, cin, cout), device="cuda") * 0.05, torch.zeros(cout, device="cuda"), fmt=fmt)
            for _ in range(3):
                K.conv(a, w, out_fmt=fmt)
            torch.cuda.synchronize()
            check(lib().mv3d_gemm_set_stamps(ptr(stamps)), "set_stamps")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            K.conv(a, w, out_fmt=fmt)
            e1.record()
            torch.cuda.synchronize()
            check(lib().mv3d_gemm_set_stamps(None), "set_stamps")
            t = stamps.cpu().tolist()
            d = [(t[i + 1] - t[i]) / mhz for i in range(7)]
            lines.append("| %s %dx%dx%d->%d | %s | %.1f | %s | %.1f |" % (name, H, W, cin, cout, "f16e5" if fmt else "bf16x3",
                         e0.elapsed_time(e1) * 1e3, " | ".join("%.1f" % v for v in d), (t[7] - t[0]) / mhz))
    text = "\n".join(lines) + "\n"
    print(text)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)


if __name__ == "__main__":
    main()
