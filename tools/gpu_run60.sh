#!/bin/bash
# final build of the round: GPU tests, trajectory modes, all modes
out=gpurun_out/r2_check_build3.txt; : > $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $out
for m in "vdp 4" "ou 12" "cir 12"; do set -- $m; SDEGPU_TRAJ_WARPS=$2 TRAJ_WARPS=$2 timeout 120 python tools/tune_traj.py $1 >> $out 2>&1; done
python tools/bench_modes.py > gpurun_out/r2_modes.jsonl 2> gpurun_out/r2_modes.err; wc -l gpurun_out/r2_modes.jsonl >> $out
cat $out
