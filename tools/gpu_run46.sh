#!/bin/bash
# k_traj: warp index through a shuffle (uniform-register bookkeeping, no R2UR loops around the bulk stores) on top of the if-chain dispatch
out=gpurun_out/r2_tune_traj_uni.txt; : > $out
run() { echo -n "$1 " >> $out; SDEGPU_TRAJ_WARPS=$3 TRAJ_WARPS=$3 timeout 120 python tools/tune_traj.py $2 >> $out 2>&1; }
tools/link_variant.sh uni || exit 1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $out
run uni vdp 4; run uni ou 12; run uni cir 12; run uni vdp 4; run uni ou 12; run uni vdp 8; run uni vdp 3
tools/link_variant.sh orig
cat $out
