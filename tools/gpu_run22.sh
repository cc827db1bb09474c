#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2_tune_stream.txt; : > $out
run() { echo -n "$1: " >> $out; timeout 300 python tools/bench_modes.py --only c3 2>/dev/null | grep suppliedB | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['name'].split('_')[-1], '%.3f ms %.0f GB/s frac %.3f;' % (d['ms'], d['gbs'], d['hbm_frac']), end=' ')
print()" >> $out; }
run orig
for v in sb_t256 sb_t128 sb_t128b2 sb_t128b3 sb_t256b1; do tools/link_variant.sh $v && run $v; done
tools/link_variant.sh orig
cat $out
