#!/bin/bash
# tools/link_variant.sh NAME [OUT]: links OUT (default sdetools_b200/libsdegpu.so) from variants/obj_NAME/*.o plus the
# regular build's objects; NAME = orig restores the regular build.
set -e
name=$1
root=$(cd "$(dirname "$0")/.." && pwd)
out=${2:-$root/sdetools_b200/libsdegpu.so}
objs=""
for o in $root/sdetools_b200/build/*.o; do
  b=$(basename $o)
  if [ "$name" != orig ] && [ -f $root/variants/obj_$name/$b ]; then objs="$objs $root/variants/obj_$name/$b"; else objs="$objs $o"; fi
done
unset CC CXX
nvcc -shared -o $out $objs -gencode arch=compute_100a,code=sm_100a -cudart static -Xlinker --exclude-libs,ALL
