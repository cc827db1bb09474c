#!/bin/bash
# Rebuilds libsdegpu.so with different k_traj block sizes and times config C3 (run on the GPU box).
out=gpurun_out/tune_traj.txt
: > $out
for t in ${TUNE_T:-512 640 768 896 1024}; do
  SDEGPU_NVCC_EXTRA="-DSDEGPU_TRAJ_THREADS=$t" python -m sdetools_b200.build --force > /dev/null 2>&1 || { echo "$t build failed" >> $out; continue; }
  python tools/bench_modes.py --only c3 2>/dev/null | grep c3_vdp_traj_philox | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('traj threads $t: %.3f ms  %.0f GB/s  hbm_frac %.3f' % (d['ms'], d['gbs'], d['hbm_frac']))" >> $out
done
python -m sdetools_b200.build --force > /dev/null 2>&1
cat $out
