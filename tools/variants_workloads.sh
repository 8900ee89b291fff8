#!/bin/bash
# frame time of some workloads for every tuning variant under rustray_b200/variants: tools/variants_workloads.sh "c2 c4_standin c5" [frames]
cd "$(dirname "$0")/.."
for f in rustray_b200/variants/*.so; do
  for w in $1; do
    echo -n "$(basename $f) $w: "
    RTX_LIB=$PWD/$f python tools/run_workload.py $w ${2:-2} 2>&1 | grep "^frame" | tail -1 | cut -c1-170
  done
done
