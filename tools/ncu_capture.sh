#!/bin/bash
# ncu captures of the hot kernels (run under gpurun; one GPU).  Outputs land in gpurun_out/.
#   tools/ncu_capture.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline --in-flight 1 --no-graph"
# 1) dominant kernel: the tap-reuse conv GEMM, full sections (skip the warm-up frames' launches)
ncu --set full --clock-control none --import-source on -k regex:conv3x3_reuse_kernel -s 200 -c 4 \
    -f -o $OUT/${TAG}_conv_full $BENCH > $OUT/${TAG}_conv_full.log 2>&1
# 2) DRAM bytes + duration of EVERY GEMM launch of one frame (49 with three views) -> roofline.traffic
for V in 3 2; do
  N=$([ $V = 3 ] && echo 49 || echo 34)
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:'conv3x3_reuse_kernel|conv_gemm_kernel' -s $((N * 8)) -c $N --csv --log-file $OUT/${TAG}_gemm_dram_v$V.csv \
      $BENCH --views $V > $OUT/${TAG}_gemm_dram_v$V.log 2>&1
done
# 3) the HBM-bound kernels: raster, ROI pool, NMS, maxpool
ncu --set full --clock-control none --import-source on -k regex:'raster_|roi_pool|nms_mask|nms_reduce|maxpool' -s 60 -c 14 \
    -f -o $OUT/${TAG}_hbm_full $BENCH > $OUT/${TAG}_hbm_full.log 2>&1
# 4) training: the backward-filter GEMM, full sections
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 70 -c 4 \
    -f -o $OUT/${TAG}_wgrad_full python bench.py --workload train --train-steps 1 --warmup 3 > $OUT/${TAG}_wgrad_full.log 2>&1
# 5) launch list of one whole inference bench run (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 500 --csv --log-file $OUT/${TAG}_launches.csv \
    $BENCH > $OUT/${TAG}_launches.log 2>&1
ls -la $OUT | tail -20
