#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "traj or golden or fuzz_shapes_fused or fused or offsets or output_options or c3" 2>&1 | tail -3
SDEGPU_TRAJ_STEPPERS=2 SDEGPU_TRAJ_GENERATORS=4 TRAJ_WARPS=2 ncu --set full --clock-control none --import-source on -k regex:k_traj -s 2 -c 1 -f -o gpurun_out/r2_prof_traj_ws python tools/tune_traj.py vdp > gpurun_out/ncu_traj_ws.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_traj_ws.log
