#!/bin/bash
# A/B of the fused ROI pool's occupancy bound (registers per thread) on ONE box: rebuilds roi_pool.cu per variant.
for n in 4 5 6 7; do
  touch mv3d_tf_b200/csrc/roi_pool.cu
  MV3D_NVCC_FLAGS=-DMV3D_ROI_MINBLOCKS=$n python -m mv3d_tf_b200.build > /dev/null 2>&1
  echo -n "minBlocks=$n: "; timeout 200 python tools/node_times.py 2>&1 | grep -E "^\| pool_5 \|"
done
touch mv3d_tf_b200/csrc/roi_pool.cu; python -m mv3d_tf_b200.build > /dev/null 2>&1
