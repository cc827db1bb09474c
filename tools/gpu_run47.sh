#!/bin/bash
# k_linear_dmma with the warp index through a shuffle (variants/obj_dmmauni) against the regular build: C4 and the dense d=128 mode
out=gpurun_out/r2_tune_dmma_uni.txt; : > $out
m() { python tools/bench_modes.py --only c4 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'frac' in d: print('$1', d['name'], '%.3f ms' % d['ms'], 'frac %.4f' % d['frac'])" >> $out; }
m orig
tools/link_variant.sh dmmauni || exit 1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dmma or linear" 2>&1 | tail -2 >> $out
m uni
tools/link_variant.sh orig
cat $out
