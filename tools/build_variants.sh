#!/bin/bash
# Build tuning variants of the CUDA library: tools/build_variants.sh "24:12 24:8 16:12 ..."  (REFILL:LEAF_BATCH)
set -e
cd "$(dirname "$0")/../rustray_b200"
mkdir -p variants
for v in $1; do
  r=${v%%:*}; l=${v##*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DRTX_REFILL=$r -DRTX_LEAF_BATCH=$l \
       -o variants/librtx_${r}_${l}.so csrc/rtx_api.cu csrc/bvh_build.cpp &
done
wait
ls -la variants
