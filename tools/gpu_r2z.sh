#!/bin/bash
# round-2 pass z (2 GPUs): sanity of the sharded path with the final library -- 2-rank parity tests, default bench line at N = 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -q > gpurun_out/r2z_pytest_sharded.log 2>&1; echo "pytest sharded: $(tail -1 gpurun_out/r2z_pytest_sharded.log)"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2z_n2_cfg2.json 2> gpurun_out/r2z_n2_cfg2.err; echo "n2 rc=$?"
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r2z_n2_cfg2.json") if l.startswith("{")][0]
print("ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "parity", d.get("parity_vs_single"), "clocks", d.get("clocks"))
PY
