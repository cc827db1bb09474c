#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
out=gpurun_out/r2_e2e_modes3.txt; : > $out
timeout 900 python tools/e2e_modes.py 131072 >> $out 2>&1
cat $out | cut -c1-420
