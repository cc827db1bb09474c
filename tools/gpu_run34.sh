#!/bin/bash
python -m pytest tests -m gpu -x -q -k "linear or tensor or tc or lin" 2>&1 | tail -2
SDEGPU_LINEAR_DMMA=0 python -m pytest tests -m gpu -x -q -k "linear or tensor or tc or lin or dmma" 2>&1 | tail -2
for v in "" "SDEGPU_LINEAR_DMMA=0" "SDEGPU_LINEAR_DMMA=0 SDEGPU_TC_NG=2" "SDEGPU_LINEAR_DMMA=0 SDEGPU_TC_NG=1"; do
echo "== $v"
env $v python tools/bench_modes.py --only linear_mid 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'dmma' in d.get('name',''): print('  ', d['name'], '%.3g path-steps/s  %.1f TFLOP/s  frac %.3f' % (d['path_steps_per_s'], d['tflops'], d['frac']))
"
done
