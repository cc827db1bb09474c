#!/bin/bash
# PhiloxPath (the path's share of the Philox rounds computed once: 17 instead of 20 wide multiplies per block), variants/obj_pp
out=gpurun_out/r2_tune_philox_path.txt; : > $out
run() { echo -n "$1 " >> $out; SDEGPU_TRAJ_WARPS=$3 TRAJ_WARPS=$3 timeout 120 python tools/tune_traj.py $2 >> $out 2>&1; }
term() { echo -n "$1 " >> $out; PROF_STEPS=10000 PROF_PATHS=10000000 timeout 120 python tools/prof_terminal.py $2 $3 $4 >> $out 2>&1; }
term orig cir heun abs; term orig ou euler identity; run orig vdp 4; run orig ou 12; run orig cir 12
tools/link_variant.sh pp || exit 1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $out
term pp cir heun abs; term pp ou euler identity; run pp vdp 4; run pp ou 12; run pp cir 12
term pp cir heun abs; term pp ou euler identity
tools/link_variant.sh orig
cat $out
