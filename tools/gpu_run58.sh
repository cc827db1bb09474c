#!/bin/bash
# k_terminal: block loop unrolled by two / three / four (variants/obj_u2, obj_u3, obj_u4)
out=gpurun_out/r2_tune_terminal_unroll34.txt; : > $out
term() { echo -n "$1 " >> $out; PROF_STEPS=10000 PROF_PATHS=10000000 timeout 120 python tools/prof_terminal.py $2 $3 $4 >> $out 2>&1; }
for v in u2 u3 u4 u2; do
  tools/link_variant.sh $v || exit 1
  term $v cir heun abs; term $v ou euler identity
done
tools/link_variant.sh orig
cat $out
