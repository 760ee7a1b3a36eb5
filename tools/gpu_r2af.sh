#!/bin/bash
# round-2 pass af: ncu --set full of the kNN full append sweep (498,196 x 128-d): stall reasons per instruction
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
KNN_SIZES=big timeout 500 ncu --set full --import-source on --clock-control none -k regex:candidates_kernel -s 2 -c 1 -f -o gpurun_out/r2af_knn_append python tools/diag/knn_timing.py > gpurun_out/r2af_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2af_ncu.log
