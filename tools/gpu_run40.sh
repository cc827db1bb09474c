#!/bin/bash
# pipelined k_traj + per-word tail + 64-bit stores for odd shifts (variants/obj_pipe2) against the regular build
out=gpurun_out/r2_tune_traj_pipe2.txt; mkdir -p gpurun_out; : > $out
run() { echo -n "$1 " >> $out; SDEGPU_TRAJ_WARPS=$3 TRAJ_WARPS=$3 timeout 120 python tools/tune_traj.py $2 >> $out 2>&1; }
term() { echo -n "$1 " >> $out; PROF_STEPS=10000 PROF_PATHS=10000000 timeout 120 python tools/prof_terminal.py $2 $3 $4 >> $out 2>&1; }
term orig cir heun abs; term orig ou euler identity
tools/link_variant.sh pipe2 || exit 1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_pipe2.log 2>&1; tail -3 gpurun_out/r2_gputests_pipe2.log >> $out
term pipe2 cir heun abs; term pipe2 ou euler identity
for w in 4 8; do run pipe2 vdp $w; done
for w in 12; do run pipe2 ou $w; run pipe2 cir $w; done
tools/link_variant.sh orig
cat $out
