#!/bin/bash
python -m pytest tests -m gpu -x -q -k "linear or tensor or tc or lin" 2>&1 | tail -3
for v in "" "SDEGPU_TC_HOT=0" "SDEGPU_TC_NG=8" "SDEGPU_TC_NG=4"; do
echo "== $v"
env $v python tools/bench_modes.py --only linear_tc 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'diag' in d.get('name',''): print('  ', d['name'], '%.3g' % d['path_steps_per_s_tc'])
"
done
