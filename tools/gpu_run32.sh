#!/bin/bash
# evidence refresh (session 4): tests, bench (own arm + reference arm), ncu launch list of the bench command, full ncu capture
# of k_terminal (summarised remotely: the .ncu-rep embeds the library's cubin and exceeds the 64 MiB return limit), all modes
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 300 gpurun_out/r2_bench_n1.json; echo
python bench.py --impl reference > gpurun_out/r2_bench_n1_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-modes > /tmp/ncu_bench.log 2>&1; echo "launch list rc=$?"
python tools/prof_terminal.py > /tmp/prof_plain.log 2>&1 && ncu --set full --clock-control none -k regex:k_terminal -s 2 -c 1 -f -o /tmp/r2_prof_terminal python tools/prof_terminal.py > /tmp/ncu_terminal.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/r2_prof_terminal.ncu-rep gpurun_out/r2_ncu_k_terminal.json > gpurun_out/r2_ncu_k_terminal.txt 2>&1; head -30 gpurun_out/r2_ncu_k_terminal.txt
python tools/bench_modes.py > gpurun_out/r2_modes.jsonl 2> gpurun_out/r2_modes.err; wc -l gpurun_out/r2_modes.jsonl; tail -2 gpurun_out/r2_modes.err
