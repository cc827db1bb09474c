#!/bin/bash
for v in "" "SDEGPU_TC_SMALL_TILE=1" "SDEGPU_TC_SMALL_TILE=1 SDEGPU_TC_NG=8" "SDEGPU_TC_NG=4"; do
echo "== $v"
env $v python tools/bench_modes.py --only linear_tc 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'diag' in d.get('name',''): print('  ', d['name'], '%.3g' % d['path_steps_per_s_tc'])
"
done
