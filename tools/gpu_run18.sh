#!/bin/bash
# 8 GPUs: config C5 at full size through the C ABI's multi-device context, two seeds; the torch.distributed arm with a second seed
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for seed in 0x5DE70005 0x5DE70006; do
  timeout 300 python tools/run_c5.py --abi-devices $N --seed $seed > gpurun_out/r2_c5_abi_n${N}_$seed.json 2> gpurun_out/r2_c5_abi_$seed.err; echo "abi rc=$?"; cat gpurun_out/r2_c5_abi_n${N}_$seed.json
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 tools/run_c5.py --seed 0x5DE70006 > gpurun_out/r2_c5_torch_n${N}_seed6.json 2> gpurun_out/r2_c5_torch.err; echo "torch rc=$?"; cat gpurun_out/r2_c5_torch_n${N}_seed6.json
timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_r_shim.py -m gpu -q -k "multi or devices" 2>&1 | tail -3
