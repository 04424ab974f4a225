#!/bin/bash
# A/B of the first-layer tcgen05 kernel's CTAs per SM on ONE box (rebuilds first_layer_tcgen05.cu per variant).
for n in 4 5 6; do
  touch mv3d_tf_b200/csrc/first_layer_tcgen05.cu
  MV3D_NVCC_FLAGS=-DMV3D_FL_MINBLOCKS=$n python -m mv3d_tf_b200.build > /dev/null 2>&1
  echo -n "CTAs/SM=$n: "; timeout 200 python tools/node_times.py 2>&1 | grep -E "^\| conv1_1_2 \|"
done
touch mv3d_tf_b200/csrc/first_layer_tcgen05.cu; python -m mv3d_tf_b200.build > /dev/null 2>&1
