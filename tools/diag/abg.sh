#!/bin/bash
# Diagnostic: the projection GEMMs alone with several builds of the library.
cd "$(dirname "$0")/../.."
for lib in "$@"; do
  echo "== $lib"
  B200GAT_LIB=$PWD/$lib timeout 100 python tools/diag/gemm_only.py 2>&1 | tail -2
done
