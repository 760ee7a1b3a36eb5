#!/bin/bash
# round-2 pass ab: kNN append pipeline -- exactness tests, then timing against the register-list kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_knn.py -x -q > gpurun_out/r2ab_knn_tests.log 2>&1; echo "knn tests rc=$? $(tail -1 gpurun_out/r2ab_knn_tests.log)"
KNN_SIZES=small timeout 200 python tools/diag/knn_timing.py > gpurun_out/r2ab_knn_timing_append.log 2>&1; echo "append rc=$?"; cat gpurun_out/r2ab_knn_timing_append.log | tail -3
B200GAT_KNN_MODE=lists KNN_SIZES=small timeout 200 python tools/diag/knn_timing.py > gpurun_out/r2ab_knn_timing_lists.log 2>&1; echo "lists rc=$?"; cat gpurun_out/r2ab_knn_timing_lists.log | tail -3
