#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/bench_modes.py --only linear_small,linear_mid,linear_tc,c4 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('  ', d.get('name'), {k:(f'{v:.3g}' if isinstance(v,float) else v) for k,v in d.items() if k in ('path_steps_per_s','path_steps_per_s_tc','tflops','frac','max_mean_err_in_se','max_cov_err_in_se')})
"
