mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_gputests_4.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2_gputests_4.log | cut -c1-300
tools/tune_traj_variants.sh
