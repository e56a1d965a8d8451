#!/bin/bash
# A/B of K2 builds: GICP_B200_LIB selects the library; knn_cov of 512 clouds x 32768 points
for lib in "$@"; do
  echo "== $lib"
  GICP_B200_LIB=$PWD/generalized-icp_b200/$lib python scripts/time_knn.py 512 0:0 2>&1 | tail -1
done
