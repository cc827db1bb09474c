#!/bin/bash
# k_linear_tc (and k_linear_dmma) with the warp index through a shuffle (variants/obj_tcuni) against the regular build
out=gpurun_out/r2_tune_tc_uni.txt; : > $out
m() { python tools/bench_modes.py --only linear_mid,linear_tc 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    v = d.get('path_steps_per_s_tc', d.get('path_steps_per_s'))
    if v and 'strict' not in d['name']: print('$1', d['name'], '%.4g' % v, ('frac %.4f' % d['frac']) if 'frac' in d else '')" >> $out; }
m orig
tools/link_variant.sh tcuni || exit 1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $out
m uni
tools/link_variant.sh orig
cat $out
