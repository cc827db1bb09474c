mkdir -p gpurun_out
TUNE_PATHS=4000000 tools/tune_variants.sh 'x_*'
cp gpurun_out/tune_variants.txt gpurun_out/r2_decomposition.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "c3_trajectory or fuzz_shapes_fused" 2>&1 | tail -3
