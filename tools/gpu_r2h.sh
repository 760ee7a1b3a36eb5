#!/bin/bash
# round-2 pass h (rerun of g with bounded output) (1 GPU): full GPU suite, bench lines of configs 1-4, ncu launch list + full-set captures of the GEMMs (config 2)
# and of the heads=4 kernels (config 3)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2h_pytest.log)"
cp gpurun_out/parity_report.json gpurun_out/r2h_parity_report.json 2>/dev/null
timeout 600 python bench.py --config 2 --steps 20 --warmup 5 > gpurun_out/r2h_cfg2.json 2> gpurun_out/r2h_cfg2.err; echo "cfg2 rc=$?"
for c in 1 4; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_cfg$c.json 2> gpurun_out/r2h_cfg$c.err; echo "cfg$c rc=$?"
done
for l in bpr bce; do
  timeout 600 python bench.py --config 3 --loss $l --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_cfg3_$l.json 2> gpurun_out/r2h_cfg3_$l.err; echo "cfg3 $l rc=$?"
done
timeout 600 python bench.py --config 2 --tier bf16 --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows > gpurun_out/r2h_cfg2_bf16.json 2> gpurun_out/r2h_cfg2_bf16.err; echo "cfg2 bf16 rc=$?"
python - <<'PY'
import json
for f in ["cfg2","cfg1","cfg4","cfg3_bpr","cfg3_bce","cfg2_bf16"]:
    try:
        d=json.load(open(f"gpurun_out/r2h_{f}.json")); print(f, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), d["breakdown_ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
B="python bench.py --config 2 --steps 2 --warmup 3 --no-cpu-baseline --no-next-rows"
$B > gpurun_out/r2h_plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2h_launches_cfg2.csv $B > gpurun_out/r2h_ncu_a.log 2>&1
echo "launch list rc=$?"
$B > gpurun_out/r2h_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"proj_kernel|proj_dw_kernel|linear128_ffma" -s 15 -c 6 -o gpurun_out/r2h_gemm_cfg2 -f $B > gpurun_out/r2h_ncu_b.log 2>&1
echo "gemm capture rc=$?"
B3="python bench.py --config 3 --steps 2 --warmup 3 --no-cpu-baseline --no-next-rows"
$B3 > gpurun_out/r2h_plain3.log 2>&1 && ncu --set full --clock-control none -k regex:"edge_fwd_kernel|edge_bwd_kernel|gemm_bf16_kernel|proj_dw_kernel" -s 24 -c 8 -o gpurun_out/r2h_cfg3 -f $B3 > gpurun_out/r2h_ncu_c.log 2>&1
echo "cfg3 capture rc=$?"
for r in r2h_gemm_cfg2 r2h_cfg3; do
  [ -f gpurun_out/$r.ncu-rep ] && ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
done
ls -la gpurun_out | grep r2g | head -40
