#!/bin/bash
# Builds libb200gat.so (sm_100a only) next to the package. Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libb200gat.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v"
mkdir -p _obj
pids=()
for f in *.cu; do
  ( $NVCC $FLAGS "$@" -c "$f" -o "_obj/${f%.cu}.o" > "_obj/${f%.cu}.log" 2>&1 || { cat "_obj/${f%.cu}.log"; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" _obj/*.o
echo "built $(realpath $OUT)"
