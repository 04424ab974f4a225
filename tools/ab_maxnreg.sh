B="python bench.py --steps 200 --warmup 5 --no-train --no-cpu-baseline --no-extras --no-other-views --no-precise-leg"
P='import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], round(d["roofline"]["frac"],3))'
echo default; $B 2>/dev/null | python -c "$P"; $B 2>/dev/null | python -c "$P"
touch mv3d_tf_b200/csrc/conv_gemm_tcgen05.cu; MV3D_NVCC_FLAGS=-DMV3D_PAIR_MAXNREG=128 python -m mv3d_tf_b200.build > /dev/null 2>&1
echo maxnreg128; $B 2>/dev/null | python -c "$P"; $B 2>/dev/null | python -c "$P"
touch mv3d_tf_b200/csrc/conv_gemm_tcgen05.cu; python -m mv3d_tf_b200.build > /dev/null 2>&1
echo default-again; $B 2>/dev/null | python -c "$P"
