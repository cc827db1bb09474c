#!/bin/bash
# Runs on the GPU box: links each variants/obj_* build in turn and times the bench workload (short), then restores
# the regular build.  Usage: tools/tune_variants.sh [pattern]   -> gpurun_out/tune_variants.txt
out=gpurun_out/tune_variants.txt
mkdir -p gpurun_out; : > $out
for d in variants/obj_${1:-*}; do
  v=${d#variants/obj_}
  tools/link_variant.sh $v || { echo "$v link failed" >> $out; continue; }
  timeout 300 python bench.py --steps 2 --warmup 1 --paths ${TUNE_PATHS:-4000000} --cpu-paths 800 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v: value %.4g  kernel_ms %.2f  mean %.9f' % (d['value'], d['roofline']['kernel_ms'], d['check']['mean']))" >> $out 2>&1
done
tools/link_variant.sh orig
cat $out
