#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "traj or golden or fuzz_shapes_fused or fused or offsets or output_options or c3" 2>&1 | tail -5
out=gpurun_out/tune_traj_tma2.txt; : > $out
run() { for m in $2; do for w in $3; do echo -n "$1 $m: " >> $out; SDEGPU_TRAJ_WARPS=$w TRAJ_WARPS=$w timeout 120 python tools/tune_traj.py $m >> $out 2>&1; done; done; }
run ch2 vdp "2 3 4"; run ch2 "ou cir" "4 6 8"
for v in tj_ch1 tj_ch4 tj_ch1r4 tj_ch2r4; do tools/link_variant.sh $v || continue
  case $v in tj_ch1) run $v vdp "3 4 6 8"; run $v "ou cir" "8";; tj_ch4) run $v vdp "2"; run $v "ou cir" "4";; tj_ch1r4) run $v vdp "6 8";; tj_ch2r4) run $v vdp "4 5";; esac
done
tools/link_variant.sh orig
cat $out
