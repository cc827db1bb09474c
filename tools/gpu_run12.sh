#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2_gputests_12.log
cat gpurun_out/r2_gputests_12.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_12.json 2> gpurun_out/r2_bench_12.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench_12.json
