#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/bench_modes.py --only linear_tc,linear_small > gpurun_out/r2_modes_tc.jsonl 2> gpurun_out/r2_modes_tc.err
cat gpurun_out/r2_modes_tc.jsonl; tail -3 gpurun_out/r2_modes_tc.err
