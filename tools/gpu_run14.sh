#!/bin/bash
mkdir -p gpurun_out
nproc; free -g | head -2
out=gpurun_out/r2_e2e_modes.txt; : > $out
for th in 1 4 16; do echo "copy threads $th" >> $out; SDEGPU_COPY_THREADS=$th timeout 600 python tools/e2e_modes.py 32768 >> $out 2>&1; done
echo "default, P=2^17" >> $out; timeout 900 python tools/e2e_modes.py 131072 >> $out 2>&1
cat $out
