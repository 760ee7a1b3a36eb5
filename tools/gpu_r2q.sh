#!/bin/bash
# round-2 pass q (1 GPU): where the time of the fp32 projection GEMM and of the kNN candidates kernel goes -- diagnostic builds
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
REPS=20 bash tools/diag/abg.sh ab/lib_g_*.so > gpurun_out/r2q_gemm_ab.log 2>&1; cat gpurun_out/r2q_gemm_ab.log
for l in ab/lib_k_*.so plotpointe-gat-recommendation_b200/libb200gat.so; do
  echo "== $l"; B200GAT_LIB=$PWD/$l KNN_SIZES=small timeout 200 python tools/diag/knn_timing.py 2>&1 | tail -2
done > gpurun_out/r2q_knn_ab.log 2>&1; cat gpurun_out/r2q_knn_ab.log
