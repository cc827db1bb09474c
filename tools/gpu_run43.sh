#!/bin/bash
# evidence refresh (session 5, final build): smoke, tests, bench (own arm + reference arm), ncu launch list of the bench command,
# full ncu captures of k_terminal and k_traj (summarised remotely), all modes
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
( time python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2>&1 | grep real; tail -c 300 gpurun_out/r2_bench_n1.json; echo
( time python bench.py --impl reference > gpurun_out/r2_bench_n1_reference_arm.json 2>/dev/null ) 2>&1 | grep real
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-modes > /tmp/ncu_bench.log 2>&1; echo "launch list rc=$?"
python tools/prof_terminal.py > /tmp/prof_plain.log 2>&1 && ncu --set full --clock-control none -k regex:k_terminal -s 2 -c 1 -f -o /tmp/r2_prof_terminal python tools/prof_terminal.py > /tmp/ncu_terminal.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/r2_prof_terminal.ncu-rep gpurun_out/r2_ncu_k_terminal.json > gpurun_out/r2_ncu_k_terminal.txt 2>&1; head -30 gpurun_out/r2_ncu_k_terminal.txt
SDEGPU_TRAJ_WARPS=4 TRAJ_WARPS=4 ncu --set full --clock-control none -k regex:k_traj -s 1 -c 1 -f -o /tmp/r2_prof_traj python tools/tune_traj.py vdp > /tmp/ncu_traj.log 2>&1; echo "ncu traj rc=$?"
python tools/ncu_summary.py /tmp/r2_prof_traj.ncu-rep gpurun_out/r2_ncu_k_traj.json > gpurun_out/r2_ncu_k_traj.txt 2>&1; tail -12 gpurun_out/r2_ncu_k_traj.txt
python tools/bench_modes.py > gpurun_out/r2_modes.jsonl 2> gpurun_out/r2_modes.err; wc -l gpurun_out/r2_modes.jsonl; tail -2 gpurun_out/r2_modes.err
