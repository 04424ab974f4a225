"""Multi-GPU correctness on hardware (SURVEY 8e, App. C last row): needs >= 2 CUDA devices (skipped otherwise; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).  The CPU-side logic of the same paths is covered
by tests/test_sharding_gloo.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_nccl_inference_bitwise_and_gradient_sum():
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_check_nccl.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("DISTCHECK ")][-1]
    out = json.loads(line[len("DISTCHECK "):])
    print(out)
    assert out["inference_bitwise_equal"] and out["inference_rows"] > 0
    assert out["grad_rel_err"] <= 1e-6, out
    assert out["grad_buckets"] >= 2 and out["grad_covered"]
