#!/bin/bash
# round-2 pass v (1 GPU): kNN with the staged (shared-memory) appends and batched list merges -- tests and timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py -q > gpurun_out/r2v_knn_tests.log 2>&1; echo "knn tests: $(tail -1 gpurun_out/r2v_knn_tests.log)"
grep -E "^FAILED|Error" gpurun_out/r2v_knn_tests.log | head -5
timeout 600 python tools/diag/knn_timing.py > gpurun_out/r2v_knn_timing.log 2>&1; tail -4 gpurun_out/r2v_knn_timing.log
B200GAT_KNN_SAMPLE=1 KNN_SIZES=small timeout 600 python tools/diag/knn_timing.py > gpurun_out/r2v_knn_timing_sampled.log 2>&1; tail -2 gpurun_out/r2v_knn_timing_sampled.log
