#!/usr/bin/env python
"""Run one 3x3 conv layer shape a few times (for `ncu -k regex:conv3x3_pair` captures of a single kernel).
    python tools/one_layer.py H W CIN COUT [f16e5|bf16x3] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mv3d_tf_b200 import kernels as K  # noqa: E402

H, W, cin, cout = (int(x) for x in sys.argv[1:5])
fmt = K.FMT_F16E5 if (len(sys.argv) < 6 or sys.argv[5] == "f16e5") else K.FMT_BF16X2
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 4
x = torch.randn((1, H, W, cin), device="cuda") * 0.5
a = K.pad_nhwc(x, precise=True, fmt=fmt)
w = K.pack_weights(torch.randn((3, 3, cin, cout), device="cuda") * 0.05, torch.zeros(cout, device="cuda"), fmt=fmt)
for _ in range(reps):
    K.conv(a, w, out_fmt=fmt)
torch.cuda.synchronize()
print("ok")
