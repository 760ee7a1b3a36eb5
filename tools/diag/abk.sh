#!/bin/bash
# Diagnostic: kNN timing (63,001 and 498,196 x 128-d) with several builds of the library.
cd "$(dirname "$0")/../.."
for lib in "$@"; do
  echo "== $lib"
  B200GAT_LIB=$PWD/$lib KNN_SIZES=small timeout 200 python tools/diag/knn_timing.py 2>&1 | tail -2
done
