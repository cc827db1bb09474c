#!/bin/bash
# k_terminal launch shapes on the final kernel: threads per SM and table replication (variants/obj_t*, obj_r16)
out=gpurun_out/r2_tune_terminal_final_shape.txt; : > $out
term() { echo -n "$1 " >> $out; PROF_STEPS=10000 PROF_PATHS=10000000 timeout 120 python tools/prof_terminal.py $2 $3 $4 >> $out 2>&1; }
for v in orig t896 t1024 r16 t896r16 orig; do
  tools/link_variant.sh $v || exit 1
  term $v cir heun abs; term $v ou euler identity
done
tools/link_variant.sh orig
cat $out
