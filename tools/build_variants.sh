#!/bin/bash
# Build tuning variants of the CUDA library: tools/build_variants.sh "name:-DFLAG=1,-DOTHER=2 name2:..."
set -e
cd "$(dirname "$0")/../rustray_b200"
mkdir -p variants
for v in $1; do
  name=${v%%:*}; flags=${v#*:}; flags=${flags//,/ }
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $flags \
       -o variants/librtx_${name}.so csrc/rtx_api.cu csrc/bvh_build.cpp &
done
wait
ls variants
