#!/bin/bash
# several path tiles per CTA in k_linear_tc (d <= 32)
python -m pytest tests -m gpu -x -q -k "linear or tensor or tc or lin" 2>&1 | tail -4
python tools/bench_modes.py --only linear_mid,linear_tc > gpurun_out/modes_tc_ng.jsonl 2> gpurun_out/modes_tc_ng.err
SDEGPU_TC_NG=1 python tools/bench_modes.py --only linear_mid,linear_tc > gpurun_out/modes_tc_ng1.jsonl 2>> gpurun_out/modes_tc_ng.err
python - <<'P'
import json
for f in ("gpurun_out/modes_tc_ng.jsonl","gpurun_out/modes_tc_ng1.jsonl"):
    print(f)
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        print("  ", d.get("name"), {k:(f"{v:.3g}" if isinstance(v,float) else v) for k,v in d.items() if k in ("path_steps_per_s","path_steps_per_s_tc","tflops","frac","ms")})
P
tail -3 gpurun_out/modes_tc_ng.err
