#!/bin/bash
# tools/build_variant.sh NAME "nvcc -D flags" [model ids...]: recompiles the given ensemble_model TUs (default: 1 = CIR)
# with extra defines into variants/obj_NAME/.  Objects (6 MB) travel to the GPU box; tools/link_variant.sh NAME links
# them there with the objects of the regular build (a linked .so is 59 MB).
set -e
name=$1; flags=$2; shift 2; models=${@:-1}
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/variants/obj_$name
for m in $models; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I $root/include \
    -DSDEGPU_TU_MODEL=$m $flags -c $root/sdetools_b200/csrc/ensemble_model.cu -o $root/variants/obj_$name/ensemble_model_$m.o
done
echo built variants/obj_$name
