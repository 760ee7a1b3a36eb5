#!/bin/bash
# round-2 pass m (N GPUs): fused projection + exchange -- parity (pytest 2-rank tests when N == 2), config 2 step time: pull / push / fused push
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29541"
if [ "$N" = 2 ]; then timeout 900 python -m pytest tests/test_gpu_sharded.py -q > gpurun_out/r2m_pytest_sharded.log 2>&1; echo "pytest sharded: $(tail -1 gpurun_out/r2m_pytest_sharded.log)"; fi
B200GAT_EXCHANGE=push timeout 300 $TR tests/sharded_check.py custom > gpurun_out/r2m_w${N}_custom_fused.log 2>&1; echo "world$N custom fused rc=$? $(grep -a SHARDED_OK gpurun_out/r2m_w${N}_custom_fused.log | head -c 200)"
run() { # tag env...
  tag=$1; shift
  env "$@" timeout 600 $TR bench.py --gpus $N --config 2 --steps 30 --warmup 5 > gpurun_out/r2m_n${N}_$tag.json 2> gpurun_out/r2m_n${N}_$tag.err; echo "n$N $tag rc=$?"
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2m_n${N}_$tag.json") if l.startswith("{")][0]
    print("  ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "comm", d["comm"]["by_kind_ms_rank0"], "kern", d["comm"]["compute_kernels_ms_per_step_max_rank"], "proj", d["breakdown_ms_per_step_rank0"].get("b200gat_project_push_f32"), d["breakdown_ms_per_step_rank0"].get("b200gat_project_bwd_push_f32"), "parity", {k:v for k,v in (d["parity_vs_single"] or {}).items() if k!="mode"})
except Exception as ex: print("  ERR", ex)
PY
}
run pull B200GAT_EXCHANGE=pull
run push B200GAT_EXCHANGE=push B200GAT_FUSED_PUSH=0
run fused B200GAT_EXCHANGE=push B200GAT_FUSED_PUSH=1
