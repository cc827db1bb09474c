#!/bin/bash
# k_traj: body loop unrolled by two (variants/obj_tu2) against the regular build
out=gpurun_out/r2_tune_traj_unroll2.txt; : > $out
run() { echo -n "$1 " >> $out; SDEGPU_TRAJ_WARPS=$3 TRAJ_WARPS=$3 timeout 120 python tools/tune_traj.py $2 >> $out 2>&1; }
run orig vdp 4; run orig ou 12; run orig cir 12
tools/link_variant.sh tu2 || exit 1
run tu2 vdp 4; run tu2 ou 12; run tu2 cir 12
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2 >> $out
tools/link_variant.sh orig
cat $out
