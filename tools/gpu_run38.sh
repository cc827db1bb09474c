#!/bin/bash
# k_traj with the generator software-pipelined one Philox block ahead (variants/obj_pipe) against the regular build:
# GPU tests on the variant, then C3 / scalar trajectories over warps per SM
out=gpurun_out/r2_tune_traj_pipe.txt; mkdir -p gpurun_out; : > $out
run() { # name model warps
  echo -n "$1 " >> $out; SDEGPU_TRAJ_WARPS=$3 TRAJ_WARPS=$3 timeout 120 python tools/tune_traj.py $2 >> $out 2>&1; }
for w in 4; do run orig vdp $w; done
run orig ou 12; run orig cir 12
tools/link_variant.sh pipe || exit 1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_pipe.log 2>&1; tail -3 gpurun_out/r2_gputests_pipe.log >> $out
for w in 2 3 4 5 6 8; do run pipe vdp $w; done
for w in 6 8 10 12; do run pipe ou $w; run pipe cir $w; done
tools/link_variant.sh orig
cat $out
