#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_19.json 2> gpurun_out/r2_bench_19.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_19_ref.json 2>/dev/null; echo "ref rc=$?"
timeout 1200 python tools/bench_modes.py > gpurun_out/r2_modes_19.jsonl 2> gpurun_out/r2_modes_19.err; echo "modes rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-modes > gpurun_out/r2_ncu_launch.log 2>&1; echo "ncu rc=$?"
head -c 600 gpurun_out/r2_bench_19.json; echo; cat gpurun_out/r2_modes_19.jsonl | cut -c1-400
