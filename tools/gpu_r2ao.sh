#!/bin/bash
# round-2 pass ao (1 GPU): bench lines of the final code for the bf16 tier and config 1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 60 python bench.py --config 2 --tier bf16 --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows > gpurun_out/r2ao_cfg2_bf16.json 2> gpurun_out/r2ao_cfg2_bf16.err; echo "bf16 rc=$?"
timeout 40 python bench.py --config 1 --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows > gpurun_out/r2ao_cfg1.json 2> gpurun_out/r2ao_cfg1.err; echo "cfg1 rc=$?"
python - <<'PY'
import json
for f in ["r2ao_cfg2_bf16","r2ao_cfg1"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3))
    except Exception as e: print(f, "ERR", e)
PY
