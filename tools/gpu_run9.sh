#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "traj or golden or fuzz_shapes_fused or fused or offsets or output_options or c3" 2>&1 | tail -5
out=gpurun_out/tune_traj_tma3.txt; : > $out
run() { for m in $2; do for w in $3; do echo -n "$1 $m: " >> $out; SDEGPU_TRAJ_WARPS=$w TRAJ_WARPS=$w timeout 120 python tools/tune_traj.py $m >> $out 2>&1; done; done; }
run ch2 vdp "3 4"
cat $out
