#!/bin/bash
# Diagnostic: config-3 (heads = 4) step with several builds of the library, one process each.
cd "$(dirname "$0")/../.."
for lib in "$@"; do
  echo "== $lib"
  B200GAT_LIB=$PWD/$lib CASES="${CASES:-h4 bf16 bpr}" timeout 200 python tools/diag/config3_timing.py 2>&1 | tail -1
done
