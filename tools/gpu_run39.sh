#!/bin/bash
# ncu capture of the pipelined k_traj (variants/obj_pipe) on C3
tools/link_variant.sh pipe || exit 1
SDEGPU_TRAJ_WARPS=4 TRAJ_WARPS=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_traj -s 1 -c 1 -f -o gpurun_out/r2_prof_traj_pipe python tools/tune_traj.py vdp > gpurun_out/ncu_traj_pipe.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_traj_pipe.log
tools/link_variant.sh orig
