#!/bin/bash
# k_terminal: block loop unrolled by two (variants/obj_u2)
out=gpurun_out/r2_tune_terminal_unroll2.txt; : > $out
term() { echo -n "$1 " >> $out; PROF_STEPS=10000 PROF_PATHS=10000000 timeout 120 python tools/prof_terminal.py $2 $3 $4 >> $out 2>&1; }
for v in orig u2 orig u2; do
  tools/link_variant.sh $v || exit 1
  term $v cir heun abs; term $v ou euler identity; term $v cir euler abs; term $v ou heun identity
done
tools/link_variant.sh u2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $out
tools/link_variant.sh orig
cat $out
