#!/bin/bash
# One profiling pass under gpurun: ncu captures (tools/ncu_capture.sh), summarised ON THE BOX so that only small tables
# travel back (gpurun merges at most 64 MiB of gpurun_out/).    gpurun --timeout 2400 -- 'bash tools/profile_round.sh r02'
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
bash tools/ncu_capture.sh $TAG ${2:-conv dram hbm launches train} > $OUT/${TAG}_capture.log 2>&1
for R in conv fc hbm; do
  if [ -f $OUT/${TAG}_${R}_full.ncu-rep ]; then
    python tools/ncu_summary.py $OUT/${TAG}_${R}_full.ncu-rep > $OUT/${TAG}_ncu_${R}_summary.md 2> $OUT/${TAG}_ncu_${R}_summary.err
  fi
done
python tools/gemm_traffic.py $TAG > $OUT/${TAG}_gemm_traffic.log 2>&1 && cp profiles/gemm_dram_traffic.json $OUT/${TAG}_gemm_dram_traffic.json
# the raw reports are large: keep only the HBM-kernel one (source-level view of the small kernels)
rm -f $OUT/${TAG}_conv_full.ncu-rep $OUT/${TAG}_fc_full.ncu-rep
du -sh $OUT; ls -la $OUT | tail -25
