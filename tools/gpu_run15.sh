#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_r_shim.py -m gpu -q -x 2>&1 | tail -8
out=gpurun_out/r2_e2e_modes2.txt; : > $out
timeout 900 python tools/e2e_modes.py 131072 >> $out 2>&1
cat $out
