#!/bin/bash
# k_linear_tc on the fast generator table (pair52): GPU tests and the linear modes
out=gpurun_out/r2_linear_tc_fastgen.txt; : > $out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_fastgen.log 2>&1; tail -5 gpurun_out/r2_gputests_fastgen.log >> $out
python tools/bench_modes.py --only linear_mid,linear_tc,c4 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print({k:(('%.4g' % v) if isinstance(v,float) else v) for k,v in d.items() if k in ('name','path_steps_per_s','path_steps_per_s_tc','tflops','tflops_tc','frac','ms')})" >> $out
cat $out
