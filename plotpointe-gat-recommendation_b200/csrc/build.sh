#!/bin/bash
# Builds libb200gat.so (sm_100a only) next to the package. Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${OUT:-../libb200gat.so}
OBJ=${OBJ:-_obj}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v"
mkdir -p $OBJ
pids=()
for f in *.cu; do
  ( $NVCC $FLAGS "$@" -c "$f" -o "$OBJ/${f%.cu}.o" > "$OBJ/${f%.cu}.log" 2>&1 || { cat "$OBJ/${f%.cu}.log"; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" $OBJ/*.o
echo "built $(realpath $OUT)"
