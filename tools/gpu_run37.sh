#!/bin/bash
# integral kernels: prefetch depth / occupancy variants (objects in variants/int, linked here against the regular build)
link() { objs=""; for o in sdetools_b200/build/*.o; do b=$(basename $o); if [ "$b" = integrals.o ] && [ "$1" != orig ]; then objs="$objs variants/int/integrals_$1.o"; else objs="$objs $o"; fi; done
  nvcc -shared -o sdetools_b200/libsdegpu.so $objs -gencode arch=compute_100a,code=sm_100a -cudart static -Xlinker --exclude-libs,ALL; }
for v in m1d1 m1d2 t384 t384d2 t192x2; do
  link $v; echo "== $v"
  python tools/bench_modes.py --only integrals 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('   %-22s %.3f ms  hbm %.3f' % (d['name'], d['ms'], d['hbm_frac']))"
done
link orig
