#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2_tune_traj_scalar.txt; : > $out
run() { for m in $2; do for w in $3; do echo -n "$1 $m: " >> $out; SDEGPU_TRAJ_WARPS=$w TRAJ_WARPS=$w timeout 120 python tools/tune_traj.py $m >> $out 2>&1; done; done; }
run orig "ou cir" "8"
tools/link_variant.sh tj_a && run tj_a "ou cir" "8 12 16"
tools/link_variant.sh tj_b && run tj_b "ou cir" "8 12"
tools/link_variant.sh orig
cat $out
