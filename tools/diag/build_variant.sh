#!/bin/bash
# Diagnostic: rebuild ONE source with extra -D flags and link it with the objects of the regular build into another library.
# usage: tools/diag/build_variant.sh <name> <source.cu> [-DFLAG=.. ...]   ->  ab/lib_<name>.so   (ab/ is git-ignored, travels with gpurun)
set -euo pipefail
cd "$(dirname "$0")/../../plotpointe-gat-recommendation_b200/csrc"
name=$1; src=$2; shift 2
mkdir -p ../../ab _obj/var
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" -c "$src" -o "_obj/var/${name}.o" > "_obj/var/${name}.log" 2>&1 || { cat "_obj/var/${name}.log"; exit 1; }
objs=$(ls _obj/*.o | grep -v "/${src%.cu}.o")
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "../../ab/lib_${name}.so" $objs "_obj/var/${name}.o"
grep -E "spill" "_obj/var/${name}.log" | grep -v "0 bytes spill stores, 0 bytes spill loads" | head -3 || true
echo "built ab/lib_${name}.so"
