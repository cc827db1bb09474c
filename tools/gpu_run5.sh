#!/bin/bash
# GPU box: trajectory-kernel parity tests, then C3 / scalar-model trajectory timings over warps per CTA
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "traj or golden or fuzz or fused or offsets or output_options or c3" > gpurun_out/r2_gputests_5.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_gputests_5.log | cut -c1-300
out=gpurun_out/tune_traj_tma.txt; : > $out
for m in vdp ou cir; do
  for w in 2 3 4 5 6 8; do echo -n "ch${TRAJ_CH:-2} $m: " >> $out; SDEGPU_TRAJ_WARPS=$w TRAJ_WARPS=$w timeout 120 python tools/tune_traj.py $m >> $out 2>&1; done
done
cat $out
