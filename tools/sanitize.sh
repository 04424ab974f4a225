#!/bin/bash
# compute-sanitizer over the kernel unit tests (race detection / sanitizers).  Run under gpurun:
#     gpurun --timeout 1500 -- 'bash tools/sanitize.sh r02'
# memcheck + racecheck over tests/test_gpu_kernels.py (raster, proposal, NMS, ROI pool, targets) and a subset of
# tests/test_gpu_gemm.py (the tcgen05 kernels hand-roll mbarrier / TMEM protocols across a CTA pair; racecheck covers
# shared-memory hazards, not the async-proxy TMA / tcgen05 traffic).  Summary lines -> gpurun_out/<tag>_sanitize.md
TAG=${1:-r02}
OUT=gpurun_out/${TAG}_sanitize.md
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
echo "# compute-sanitizer summary (${TAG})" > $OUT
echo >> $OUT
run() {  # tool, label, time limit, pytest args...
  local tool=$1 label=$2 limit=$3; shift 3
  local log=gpurun_out/${TAG}_sanitize_${tool}_${label}.log
  timeout $limit $CS --tool $tool --print-limit 20 --error-exitcode 0 python -m pytest -x -q -m gpu -p no:cacheprovider "$@" > $log 2>&1
  local rc=$?
  echo "## $tool: $label (rc=$rc, limit ${limit}s)" >> $OUT
  echo '```' >> $OUT
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $log | tail -8 >> $OUT
  grep -E "Invalid|Race reported|hazard" $log | sort | uniq -c | head -20 >> $OUT
  echo '```' >> $OUT
}
SET=${2:-all}   # all | new (only the kernels added late in round 2: cluster NMS, first-layer tcgen05, f16e5 ROI/fc operands)
run memcheck edge 600 tests/test_gpu_edge_cases.py -k "nms or roi_pool"
run racecheck edge 600 tests/test_gpu_edge_cases.py -k "nms_lazy or emits_f16e5 or pad_operand"
run memcheck first_layer 300 tests/test_gpu_gemm.py -k "first_layer_direct or f16e5_operands"
run racecheck first_layer 300 tests/test_gpu_gemm.py -k "first_layer_direct and not 1242"
if [ "$SET" = all ]; then
run memcheck kernels 600 tests/test_gpu_kernels.py
run memcheck gemm 600 tests/test_gpu_gemm.py -k "not full and not big"
run racecheck kernels 600 tests/test_gpu_kernels.py -k "nms or roi or raster"
run racecheck gemm 420 tests/test_gpu_gemm.py -k "pair or f16e5"
fi
cat $OUT
