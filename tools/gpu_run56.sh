#!/bin/bash
# k_linear_tc: 64 < d <= 128 on 32-path tiles with 384 threads per CTA (SDEGPU_TC_MID384=1), variants/obj_m384
out=gpurun_out/r2_tune_tc_mid384.txt; : > $out
m() { SDEGPU_TC_MID384=$2 python tools/bench_modes.py --only linear_mid 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'frac' in d: print('$1', d['name'], '%.4g' % d['path_steps_per_s'], 'frac %.4f' % d['frac'])" >> $out; }
tools/link_variant.sh m384 || exit 1
m tiles64x256thr 0
m tiles32x384thr 1
SDEGPU_TC_MID384=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "linear or tensor or dmma or tc" 2>&1 | tail -2 >> $out
tools/link_variant.sh orig
cat $out
