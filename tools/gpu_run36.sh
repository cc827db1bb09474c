#!/bin/bash
# 8 GPUs, final build: config C5 at full size through the C ABI's multi-GPU context (one process) and the bench under torchrun
N=${1:-8}
timeout 300 python tools/run_c5.py --abi-devices $N --seed 0x5DE70007 > gpurun_out/r2_c5_full_size_abi_n${N}_seed7.json 2> gpurun_out/r2_c5_abi_seed7.err; echo "abi rc=$?"; cat gpurun_out/r2_c5_full_size_abi_n${N}_seed7.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; echo "bench rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1]); print(d['n_gpus'], '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['ms_per_step'], d['clocks'])"
