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
# the same SASS check the product build gets: every warp vote of the traversal kernels must carry its BRA.DIV / WARPSYNC guard
cd .. && python -c "
import glob, __graft_entry__ as g
for f in sorted(glob.glob('rustray_b200/variants/*.so')):
    g.check_vote_convergence(f); print('votes guarded:', f)"
