#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -k "tensor or linear or dmma" 2>&1 | tail -12
timeout 900 python tools/bench_modes.py --only linear_tc 2>&1 | grep -E "d512|d256_heun" | cut -c1-330
