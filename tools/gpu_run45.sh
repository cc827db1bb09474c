#!/bin/bash
# k_traj: row-phase dispatch by two compare-and-branch levels instead of a jump table (variants/obj_ifs); the new tail-draw test
out=gpurun_out/r2_tune_traj_ifs.txt; : > $out
run() { echo -n "$1 " >> $out; SDEGPU_TRAJ_WARPS=$3 TRAJ_WARPS=$3 timeout 120 python tools/tune_traj.py $2 >> $out 2>&1; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tail_draws or fuzz_shapes_fused or output_options" 2>&1 | tail -3 >> $out
run build vdp 4; run build ou 12
tools/link_variant.sh ifs || exit 1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2 >> $out
run ifs vdp 4; run ifs ou 12; run ifs vdp 4; run ifs ou 12
tools/link_variant.sh orig
cat $out
