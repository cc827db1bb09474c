#!/bin/bash
# final build check: GPU tests, linear modes, C4, trajectories
out=gpurun_out/r2_check_build2.txt; : > $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $out
python tools/bench_modes.py --only linear_mid,linear_tc,c4,c3 2>/dev/null > gpurun_out/r2_modes_partial.jsonl
python -c "
import json
for l in open('gpurun_out/r2_modes_partial.jsonl'):
    d=json.loads(l); v = d.get('path_steps_per_s_tc', d.get('path_steps_per_s'))
    if v and 'strict_cuda' not in d['name']: print(d['name'], '%.4g' % v, ('frac %.4f' % d.get('frac', d.get('hbm_frac'))) if ('frac' in d or 'hbm_frac' in d) else '')" >> $out
cat $out
