#!/bin/bash
# round-2 pass ac: per-kernel times of the kNN append pipeline (ncu, time only)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
KNN_SIZES=small timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ac_knn_launches.csv python tools/diag/knn_timing.py > gpurun_out/r2ac_ncu.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2ac_knn_launches.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
for r in rows[1:]:
    print(r[ki][:90], r[vi])
PY
