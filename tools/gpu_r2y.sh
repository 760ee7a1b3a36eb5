#!/bin/bash
# round-2 pass y (1 GPU): records of the final code -- driver sequence (smoke, GPU suite, default bench line, reference arm), bench lines of
# configs 1 / 3 / 4, launch list of the config-2 step, ncu full-set capture of the kNN candidates kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
bash tools/gpu_final.sh
for c in 1 4; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2y_cfg$c.json 2> gpurun_out/r2y_cfg$c.err; echo "cfg$c rc=$?"
done
timeout 600 python bench.py --config 3 --loss bpr --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2y_cfg3_bpr.json 2> gpurun_out/r2y_cfg3_bpr.err; echo "cfg3 rc=$?"
timeout 600 python bench.py --config 2 --tier bf16 --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows > gpurun_out/r2y_cfg2_bf16.json 2> gpurun_out/r2y_cfg2_bf16.err; echo "cfg2 bf16 rc=$?"
python - <<'PY'
import json
for f in ["cfg1","cfg4","cfg3_bpr","cfg2_bf16"]:
    try:
        d=json.load(open(f"gpurun_out/r2y_{f}.json")); print(f, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3))
    except Exception as e: print(f, "ERR", e)
PY
B="python bench.py --config 2 --steps 2 --warmup 3 --no-cpu-baseline --no-next-rows"
$B > gpurun_out/r2y_plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2y_launches_cfg2.csv $B > gpurun_out/r2y_ncu_a.log 2>&1
echo "launch list rc=$?"
K="env KNN_SIZES=one python tools/diag/knn_timing.py"
$K > gpurun_out/r2y_knn_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:candidates_kernel -c 2 -o gpurun_out/r2y_knn -f $K > gpurun_out/r2y_ncu_k.log 2>&1
echo "knn capture rc=$?"
[ -f gpurun_out/r2y_knn.ncu-rep ] && ncu -i gpurun_out/r2y_knn.ncu-rep --page raw --csv > gpurun_out/r2y_knn_raw.csv 2>/dev/null
ls -la gpurun_out | grep r2y | head -30
