#!/bin/bash
# what the driver runs at round end, on one GPU: smoke(), the GPU suite, the default bench line and the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$? $(tail -2 gpurun_out/final_smoke.log | tr '\n' ' ')"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/final_pytest.log)"
cp gpurun_out/parity_report.json gpurun_out/final_parity_report.json 2>/dev/null
T1=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$? in $(( $(date +%s) - T1 )) s"
T2=$(date +%s)
timeout 1200 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "reference rc=$? in $(( $(date +%s) - T2 )) s"
python - <<'PY'
import json
d=json.load(open("gpurun_out/final_bench.json")); r=json.load(open("gpurun_out/final_ref.json"))
print("ours ms", round(d["ms_per_step"],3), "value %.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("dram_frac"), "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
print("ref value %.4g"%r["value"], "ms", round(r["ms_per_step"],1), r["cpu_baseline"]["kind"], r["cpu_baseline"]["cores"], "ratio e2e", d["e2e"]["value"]/r["value"])
print(d["breakdown_ms_per_step"])
PY
echo "total $(( $(date +%s) - T0 )) s"
