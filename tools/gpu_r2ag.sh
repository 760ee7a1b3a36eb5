#!/bin/bash
# round-2 pass ag: kNN append pipeline, 16 vs 8 epilogue warps in the append sweeps (tests + timing each)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for lib in plotpointe-gat-recommendation_b200/libb200gat.so ab/lib_k8w.so; do
  tag=$(basename $lib .so)
  B200GAT_LIB=$PWD/$lib timeout 420 python -m pytest tests/test_gpu_knn.py -x -q > gpurun_out/r2ag_tests_$tag.log 2>&1; echo "$tag tests rc=$? $(tail -1 gpurun_out/r2ag_tests_$tag.log)"
  B200GAT_LIB=$PWD/$lib KNN_SIZES=small timeout 200 python tools/diag/knn_timing.py > gpurun_out/r2ag_timing_$tag.log 2>&1; tail -2 gpurun_out/r2ag_timing_$tag.log
done
