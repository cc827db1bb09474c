#!/bin/bash
# staging-copy experiment: streaming stores on/off, copy threads, against the box's pinned PCIe rates
set -x
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|^CPU\(s\)"; free -g | head -2
python tools/e2e_modes.py 131072 > gpurun_out/e2e_nt1.txt 2>&1
SDEGPU_COPY_NT=0 E2E_PCIE_PEAK=0 python tools/e2e_modes.py 131072 > gpurun_out/e2e_nt0.txt 2>&1
SDEGPU_COPY_THREADS=8 E2E_PCIE_PEAK=0 python tools/e2e_modes.py 131072 > gpurun_out/e2e_nt1_t8.txt 2>&1
SDEGPU_COPY_THREADS=32 E2E_PCIE_PEAK=0 python tools/e2e_modes.py 131072 > gpurun_out/e2e_nt1_t32.txt 2>&1
tail -n 5 gpurun_out/e2e_nt*.txt
