#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "fuzz_shapes_fused" 2>&1 | grep -B30 "AssertionError" | head -60
w=4
SDEGPU_TRAJ_WARPS=$w TRAJ_WARPS=$w ncu --set full --clock-control none --import-source on -k regex:k_traj -s 2 -c 1 -f -o gpurun_out/r2_prof_traj_b python tools/tune_traj.py vdp > gpurun_out/ncu_traj_b.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_traj_b.log
