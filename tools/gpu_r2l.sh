#!/bin/bash
# round-2 pass l (N GPUs): push vs pull gathers -- parity, then config 2 step time with each
N=${1:-4}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29541"
for mode in push pull; do
  B200GAT_EXCHANGE=$mode timeout 300 $TR tests/sharded_check.py pyg > gpurun_out/r2l_w${N}_$mode.log 2>&1; echo "world$N $mode rc=$? $(grep -a SHARDED_OK gpurun_out/r2l_w${N}_$mode.log | head -c 200)"
  for tier in f32 bf16; do
    B200GAT_EXCHANGE=$mode timeout 600 $TR bench.py --gpus $N --config 2 --tier $tier --steps 30 --warmup 5 > gpurun_out/r2l_n${N}_${mode}_$tier.json 2> gpurun_out/r2l_n${N}_${mode}_$tier.err; echo "n$N $mode $tier rc=$?"
    python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2l_n${N}_${mode}_$tier.json") if l.startswith("{")][0]
    print("  ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "comm", d["comm"]["by_kind_ms_rank0"], "GB/s", d["comm"]["pull_gbs_rank0"], "kern", d["comm"]["compute_kernels_ms_per_step_max_rank"], "parity", {k:v for k,v in (d["parity_vs_single"] or {}).items() if k!="mode"})
except Exception as ex: print("  ERR", ex)
PY
  done
done
