#!/bin/bash
# round-2 pass al (1 GPU): edge kernels with cp.async rings as the default -- smoke, GPU suite, default bench line, ncu --set full of the
# edge kernels (traffic.json), launch list of the step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2al_smoke.log 2>&1; echo "smoke rc=$? $(tail -2 gpurun_out/r2al_smoke.log | tr '\n' ' ')"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2al_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2al_pytest.log)"
cp gpurun_out/parity_report.json gpurun_out/r2al_parity_report.json 2>/dev/null
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2al_bench.json 2> gpurun_out/r2al_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2al_bench.json"))
print("ours ms", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("dram_frac"), d["roofline"].get("other_edge_kernels"))
print(d["breakdown_ms_per_step"])
PY
B="python bench.py --config 2 --steps 2 --warmup 3 --no-cpu-baseline --no-next-rows"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_ -c 4 -f -o gpurun_out/r2al_edge $B > gpurun_out/r2al_ncu_edge.log 2>&1
echo "edge capture rc=$?"
[ -f gpurun_out/r2al_edge.ncu-rep ] && ncu -i gpurun_out/r2al_edge.ncu-rep --page raw --csv > gpurun_out/r2al_edge_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2al_edge_raw.csv "r2al" | tail -5
