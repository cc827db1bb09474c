#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "traj or golden or fuzz_shapes_fused or fused or offsets or output_options or c3" 2>&1 | tail -5
out=gpurun_out/tune_traj_ws.txt; : > $out
run() { for g in $3; do echo -n "$1 S=$2 G=$g $4: " >> $out; SDEGPU_TRAJ_STEPPERS=$2 SDEGPU_TRAJ_GENERATORS=$g TRAJ_WARPS=$2 timeout 120 python tools/tune_traj.py $4 >> $out 2>&1; done; }
run ws 2 "4 6 8 10" vdp
run ws 3 "6 9" vdp
run ws 4 "8" vdp
run ws 4 "4 8" ou
run ws 4 "4 8" cir
run ws 8 "8" ou
cat $out
