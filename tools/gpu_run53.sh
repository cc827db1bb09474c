#!/bin/bash
# k_linear_tc: 16-path tiles (NBT = 2) for 16 < d <= 32, one warp per tile (variants/obj_nbt2; SDEGPU_TC_NBT2=0 keeps the 32-path tiles)
out=gpurun_out/r2_tune_tc_nbt2.txt; : > $out
m() { SDEGPU_TC_NBT2=$2 python tools/bench_modes.py --only linear_tc 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'd32' in d['name'] or 'd16' in d['name']: print('$1', d['name'], '%.4g' % d['path_steps_per_s_tc'])" >> $out; }
tools/link_variant.sh nbt2 || exit 1
m tiles32 0
m tiles16 1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $out
tools/link_variant.sh orig
cat $out
