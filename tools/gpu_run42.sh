#!/bin/bash
# regular build after the generator/trajectory changes: GPU tests, C2/C5-share rates, trajectory modes
out=gpurun_out/r2_check_build.txt; : > $out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_final.log 2>&1; tail -3 gpurun_out/r2_gputests_final.log >> $out
term() { echo -n "$1 " >> $out; PROF_STEPS=10000 PROF_PATHS=10000000 timeout 120 python tools/prof_terminal.py $2 $3 $4 >> $out 2>&1; }
term build cir heun abs; term build ou euler identity
for m in vdp ou cir; do timeout 120 python tools/tune_traj.py $m >> $out 2>&1; done
cat $out
