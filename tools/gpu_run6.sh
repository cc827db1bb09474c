#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "fuzz_shapes_fused" 2>&1 | tail -5
for w in 4 8; do
SDEGPU_TRAJ_WARPS=$w TRAJ_WARPS=$w ncu --set full --clock-control none --import-source on -k regex:k_traj -s 2 -c 1 -f -o gpurun_out/r2_prof_traj_w$w python tools/tune_traj.py vdp > gpurun_out/ncu_traj_w$w.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_traj_w$w.log
done
