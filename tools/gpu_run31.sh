#!/bin/bash
# two devices: the multi-device tests, smoke(), the bench under torchrun, the reference arm
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/s4_bench_n2.json 2> gpurun_out/s4_bench_n2.err; tail -c 600 gpurun_out/s4_bench_n2.json; tail -2 gpurun_out/s4_bench_n2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s4_bench_ref.json 2>/dev/null; head -c 700 gpurun_out/s4_bench_ref.json
