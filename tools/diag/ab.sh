#!/bin/bash
# Diagnostic: time the config-2 step with several builds of the library (ab/lib*.so), one process each.
cd "$(dirname "$0")/../.."
for lib in "$@"; do
  echo "== $lib"
  B200GAT_LIB=$PWD/$lib TIERS=${TIERS:-f32} timeout 200 python tools/diag/tier_timing.py 2>&1 | tail -2
done
