#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2_gputests_16.log
cat gpurun_out/r2_gputests_16.log
