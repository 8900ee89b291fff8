#!/bin/bash
cd "$(dirname "$0")/.."
for f in rustray_b200/librtx_b200.so rustray_b200/variants/*.so; do
  for scene in room_spheres kbert c1_spheres; do
    QUIET=1 RTX_LIB=$PWD/$f python tools/gpu_debug4.py 400 225 0 $scene 2>&1 | tail -1
  done
done
