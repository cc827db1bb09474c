mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_gputests_2.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_gputests_2.log
python tools/prof_terminal.py > gpurun_out/prof_plain.log 2>&1 && cat gpurun_out/prof_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:k_terminal -s 2 -c 1 -f -o gpurun_out/r2_prof_terminal python tools/prof_terminal.py > gpurun_out/ncu_terminal.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_terminal.log
