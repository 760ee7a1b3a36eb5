#!/bin/bash
# round-2 pass p (1 GPU): kNN with the sampled admission threshold: tests, timing A/B, ncu of the candidates kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py -q > gpurun_out/r2p_knn_tests.log 2>&1; echo "knn tests: $(tail -1 gpurun_out/r2p_knn_tests.log)"
grep -E "^FAILED|Error" gpurun_out/r2p_knn_tests.log | head -5
timeout 600 python tools/diag/knn_timing.py > gpurun_out/r2p_knn_timing_sampled.log 2>&1; cat gpurun_out/r2p_knn_timing_sampled.log | tail -4
B200GAT_KNN_NO_SAMPLE=1 timeout 600 python tools/diag/knn_timing.py > gpurun_out/r2p_knn_timing_plain.log 2>&1; cat gpurun_out/r2p_knn_timing_plain.log | tail -4
