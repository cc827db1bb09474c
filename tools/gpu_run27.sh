#!/bin/bash
# duplex staging (upload of chunk k alongside download of chunk k-1) and the copy-thread default
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/e2e_modes.py 131072 > gpurun_out/e2e_duplex.txt 2>&1
SDEGPU_NO_DUPLEX=1 E2E_PCIE_PEAK=0 python tools/e2e_modes.py 131072 > gpurun_out/e2e_noduplex.txt 2>&1
tail -n 4 gpurun_out/e2e_duplex.txt gpurun_out/e2e_noduplex.txt | cut -c1-420
