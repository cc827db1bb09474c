#!/bin/bash
# round-2 ncu summaries of the HBM-bound kernels beside k_traj: k_stream (C3 with a supplied B) and k_integral / k_rvbm
ncu --set full --clock-control none -k regex:k_stream -c 2 -f -o /tmp/r2_prof_stream python tools/bench_modes.py --only c3 > /tmp/ncu_stream.log 2>&1; echo "ncu stream rc=$?"
python tools/ncu_summary.py /tmp/r2_prof_stream.ncu-rep gpurun_out/r2_ncu_k_stream.json > gpurun_out/r2_ncu_k_stream.txt 2>&1
ncu --set full --clock-control none -k regex:"k_integral|k_rvbm" -c 4 -f -o /tmp/r2_prof_int python tools/bench_modes.py --only integrals,rvbm > /tmp/ncu_int.log 2>&1; echo "ncu integrals rc=$?"
python tools/ncu_summary.py /tmp/r2_prof_int.ncu-rep gpurun_out/r2_ncu_k_integral.json > gpurun_out/r2_ncu_k_integral.txt 2>&1
grep -E "^==|duration|dram__bytes|stalls" gpurun_out/r2_ncu_k_stream.txt gpurun_out/r2_ncu_k_integral.txt
