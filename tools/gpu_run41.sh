#!/bin/bash
# k_terminal: rare path out of line (variants/obj_rare), table address by IMAD.HI on the FMA pipe (variants/obj_imad)
out=gpurun_out/r2_tune_terminal_imad.txt; mkdir -p gpurun_out; : > $out
term() { echo -n "$1 " >> $out; PROF_STEPS=10000 PROF_PATHS=10000000 timeout 120 python tools/prof_terminal.py $2 $3 $4 >> $out 2>&1; }
for v in orig rare imad orig rare imad; do
  tools/link_variant.sh $v || exit 1
  term $v cir heun abs; term $v ou euler identity
done
tools/link_variant.sh rare
echo -n "rare " >> $out; SDEGPU_TRAJ_WARPS=4 TRAJ_WARPS=4 timeout 120 python tools/tune_traj.py vdp >> $out 2>&1
tools/link_variant.sh orig
cat $out
