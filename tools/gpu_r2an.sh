#!/bin/bash
# round-2 pass an (1 GPU): final code (cp.async rings in every edge kernel) -- GPU suite, default bench line, config-3 bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2an_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2an_pytest.log)"
cp gpurun_out/parity_report.json gpurun_out/r2an_parity_report.json 2>/dev/null
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2an_bench.json 2> gpurun_out/r2an_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --config 3 --loss bpr --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows > gpurun_out/r2an_cfg3_bpr.json 2> gpurun_out/r2an_cfg3.err; echo "cfg3 rc=$?"
python - <<'PY'
import json
for f in ["r2an_bench","r2an_cfg3_bpr"]:
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, "ms", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"].get("dram_frac"))
    print("  ", d["breakdown_ms_per_step"])
PY
