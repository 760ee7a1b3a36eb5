#!/bin/bash
# round-2 pass ae: kNN append pipeline -- exactness tests, plain timing, per-kernel times (ncu, time only)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_knn.py -x -q > gpurun_out/r2ae_knn_tests.log 2>&1; echo "knn tests rc=$? $(tail -1 gpurun_out/r2ae_knn_tests.log)"
KNN_SIZES=${KNN_SIZES:-small} timeout 200 python tools/diag/knn_timing.py > gpurun_out/r2ae_knn_timing.log 2>&1; echo "timing rc=$?"; tail -4 gpurun_out/r2ae_knn_timing.log
KNN_SIZES=small timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ae_knn_launches.csv python tools/diag/knn_timing.py > gpurun_out/r2ae_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2ae_knn_launches.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
out=[(r[ki][:70], float(r[vi].replace(",",""))/1e6) for r in rows[1:] if "b200gat" in r[ki]]
for k,v in out[len(out)//2+len(out)//4:]: print(f"{v:8.3f} ms  {k}")
PY
