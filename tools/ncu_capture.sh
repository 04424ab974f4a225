#!/bin/bash
# ncu captures of the hot kernels (run under gpurun; one GPU).  Outputs land in gpurun_out/.
#   tools/ncu_capture.sh <tag> [steps: conv dram hbm wgrad launches train]
set -u
TAG=${1:-r02}
shift
STEPS=${*:-conv dram hbm wgrad launches train}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline --in-flight 1 --no-graph --mode mixed --no-extras --no-other-views --no-precise-leg"
GEMMS='conv3x3_pair_kernel|conv3x3_reuse_kernel|conv_gemm_kernel|fc_swapped_pair_kernel'
for S in $STEPS; do
case $S in
conv)  # dominant kernels: the CTA-pair tap-reuse conv (all widths; 26 launches per 2-view frame) + the swapped fc kernel, full sections
  ncu --set full --clock-control none --import-source on -k regex:conv3x3_pair_kernel -s 104 -c 26 \
      -f -o $OUT/${TAG}_conv_full $BENCH > $OUT/${TAG}_conv_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:fc_swapped_pair_kernel -s 16 -c 2 \
      -f -o $OUT/${TAG}_fc_full $BENCH > $OUT/${TAG}_fc_full.log 2>&1 ;;
dram)  # DRAM bytes + duration of EVERY GEMM launch of one frame -> roofline.traffic
  for V in 3 2; do
    N=$([ $V = 3 ] && echo 46 || echo 32)
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:"$GEMMS" -s $((N * 5)) -c $N --csv --log-file $OUT/${TAG}_gemm_dram_v$V.csv \
        $BENCH --views $V > $OUT/${TAG}_gemm_dram_v$V.log 2>&1
  done ;;
hbm)   # the HBM-/latency-bound kernels: raster, first image layer, ROI pool, NMS, proposal chain, remaining max-pools
  ncu --set full --clock-control none --import-source on -k regex:'raster_|roi_pool|nms_|maxpool|small_cin|proposal_' -s 60 -c 15 \
      -f -o $OUT/${TAG}_hbm_full $BENCH > $OUT/${TAG}_hbm_full.log 2>&1 ;;
wgrad) # training: the backward-filter GEMM, full sections
  ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 70 -c 4 \
      -f -o $OUT/${TAG}_wgrad_full python bench.py --workload train --train-steps 1 --warmup 3 > $OUT/${TAG}_wgrad_full.log 2>&1 ;;
launches) # launch list of one whole inference bench run (shares, not absolutes)
  ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 300 --csv --log-file $OUT/${TAG}_ncu_launches_infer.csv \
      $BENCH --steps 4 > $OUT/${TAG}_launches.log 2>&1 ;;
train) # launch list of two train steps
  ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 600 --csv --log-file $OUT/${TAG}_ncu_launches_train.csv \
      python bench.py --workload train --train-steps 2 --warmup 3 > $OUT/${TAG}_train_launches.log 2>&1 ;;
esac
done
ls -la $OUT | tail -20
