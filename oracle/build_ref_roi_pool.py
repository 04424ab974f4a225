"""TEST INFRASTRUCTURE ONLY.  Recipe: compile the reference's lib/roi_pooling_layer/roi_pooling_op.cc -- unmodified,
from where it lies under /root/reference -- against the stand-in TensorFlow headers in oracle/tf_stub/ plus the C driver
oracle/ref_roi_pool_driver.cc, into oracle/_ref/libref_roi_pool.so (git-ignored; travels to the GPU box).
    python -m oracle.build_ref_roi_pool
No reference source is copied into the repository; without /root/reference the recipe is a no-op."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/lib/roi_pooling_layer"
OUT = os.path.join(HERE, "_ref", "libref_roi_pool.so")


def build(force=False):
    src = os.path.join(REF, "roi_pooling_op.cc")
    if not os.path.exists(src):
        return OUT if os.path.exists(OUT) else None
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [src, os.path.join(HERE, "ref_roi_pool_driver.cc"), os.path.join(HERE, "tf_stub", "tensorflow", "core", "framework", "op_kernel.h")]
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++14", "-shared", "-fPIC", "-w", "-I", os.path.join(HERE, "tf_stub"), "-I", REF,
                        src, os.path.join(HERE, "ref_roi_pool_driver.cc"), "-o", OUT], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
