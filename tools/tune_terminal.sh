#!/bin/bash
# Rebuilds libsdegpu.so with different k_terminal launch shapes and times the bench workload (run on the GPU box).
out=gpurun_out/tune_terminal.txt
: > $out
for cfg in ${TUNE_CFGS:-"896 1" "768 1" "640 1" "1024 1" "448 2" "384 2" "512 1"}; do
  set -- ${cfg//,/ }
  SDEGPU_NVCC_EXTRA="-DSDEGPU_TERMINAL_THREADS=$1 -DSDEGPU_TERMINAL_MIN_BLOCKS=$2 -DSDEGPU_TERMINAL_NP=${3:-1}" python -m sdetools_b200.build --force > /dev/null 2>&1 || { echo "$cfg build failed" >> $out; continue; }
  python bench.py --steps 2 --warmup 1 --paths 4000000 --cpu-paths 800 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('threads $1 minblocks $2 np ${3:-1}: value %.4g  kernel_ms %.2f frac %.3f' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))" >> $out
done
python -m sdetools_b200.build --force > /dev/null 2>&1
cat $out
