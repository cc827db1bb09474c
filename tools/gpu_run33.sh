#!/bin/bash
# generator evaluation with shared segment bits in the 16-normal bodies of k_traj: regular build vs variants/obj_multi
run() { python tools/bench_modes.py --only c3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'ms' in d: print('  ', d['name'], '%.3f ms' % d['ms'], 'hbm %.3f' % d.get('hbm_frac', 0))
"; }
echo "== regular"; run
tools/link_variant.sh multi && { echo "== multi"; run; python -m pytest tests -m gpu -x -q -k "traj or guard or write_nothing or c3 or models" 2>&1 | tail -2; }
tools/link_variant.sh orig
