#!/bin/bash
# ncu captures of the hot kernels (run under gpurun; one GPU).  Outputs land in gpurun_out/.
#   tools/ncu_capture.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
# 1) dominant kernel: the tap-reuse conv GEMM (skip the warm-up frames' launches)
ncu --set full --clock-control none --import-source on -k regex:conv3x3_reuse_kernel -s 60 -c 4 \
    -f -o $OUT/${TAG}_conv_full $BENCH > $OUT/${TAG}_conv_full.log 2>&1
# 2) the HBM-bound kernels: raster, ROI pool, NMS, maxpool
ncu --set full --clock-control none --import-source on -k regex:'raster_tile|roi_pool|nms_mask|nms_reduce|maxpool' -s 30 -c 10 \
    -f -o $OUT/${TAG}_hbm_full $BENCH > $OUT/${TAG}_hbm_full.log 2>&1
# 3) launch list of one whole bench run (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    $BENCH > $OUT/${TAG}_launches.log 2>&1
ls -la $OUT
