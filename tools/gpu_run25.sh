#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_linear_tc.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_linear_tc -c 2 -f -o gpurun_out/r2_prof_linear_tc python tools/prof_linear_tc.py > gpurun_out/ncu_linear_tc.log 2>&1; echo "ncu rc=$?"
SDEGPU_TRAJ_WARPS=12 TRAJ_WARPS=12 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_traj -s 1 -c 1 -f -o gpurun_out/r2_prof_traj_cir python tools/tune_traj.py cir > gpurun_out/ncu_traj_cir.log 2>&1; echo "ncu rc=$?"
