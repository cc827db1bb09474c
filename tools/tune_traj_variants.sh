#!/bin/bash
# GPU box: for each variants/obj_tj_* build (name tj_tw<TW>w<WARPS>...), link and time C3 (+ scalar models)
out=gpurun_out/tune_traj.txt; mkdir -p gpurun_out; : > $out
for d in variants/obj_tj_*; do
  v=${d#variants/obj_}
  w=$(echo $v | sed -E 's/.*w([0-9]+).*/\1/')
  tools/link_variant.sh $v || { echo "$v link failed" >> $out; continue; }
  for m in ${TRAJ_MODELS:-vdp}; do echo -n "$v: " >> $out; TRAJ_WARPS=$w timeout 120 python tools/tune_traj.py $m >> $out 2>&1; done
done
tools/link_variant.sh orig
cat $out
