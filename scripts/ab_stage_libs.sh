#!/bin/bash
# A/B of whole-library builds on the per-stage times of 512 bench pairs: bash scripts/ab_stage_libs.sh libA.so libB.so ...
for lib in "$@"; do
  echo "== $lib"
  GICP_B200_LIB=$PWD/generalized-icp_b200/$lib python scripts/stage_times.py 512 3 2>&1 | tail -7 | tr '\n' ';'; echo
done
