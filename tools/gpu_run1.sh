mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_1.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_gputests_1.log
timeout 300 python bench.py --steps 3 --warmup 2 > gpurun_out/r2_bench_1.json 2> gpurun_out/r2_bench_1.err; echo "bench rc=$?"; cat gpurun_out/r2_bench_1.json | cut -c1-600
tools/tune_variants.sh
