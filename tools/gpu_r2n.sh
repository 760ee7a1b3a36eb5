#!/bin/bash
# round-2 pass n (N GPUs): final scaling records of the default exchange (pushes, fused into the GEMM epilogues where the shape allows)
N=${1:-8}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29541"
for k in custom; do
  timeout 300 $TR tests/sharded_check.py $k > gpurun_out/r2n_w${N}_$k.log 2>&1; echo "world$N $k rc=$? $(grep -a SHARDED_OK gpurun_out/r2n_w${N}_$k.log | head -c 300)"
done
run() { # tag, args...
  tag=$1; shift
  timeout 600 $TR bench.py --gpus $N "$@" > gpurun_out/r2n_n${N}_$tag.json 2> gpurun_out/r2n_n${N}_$tag.err; echo "n$N $tag rc=$?"
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/r2n_n${N}_$tag.json") if l.startswith("{")][0]
    print("  ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "comm", d["comm"]["by_kind_ms_rank0"], "kern", d["comm"]["compute_kernels_ms_per_step_max_rank"], "parity", {k:v for k,v in (d["parity_vs_single"] or {}).items() if k!="mode"})
    print("  ", d["breakdown_ms_per_step_rank0"])
except Exception as ex: print("  ERR", ex)
PY
}
run cfg2 --config 2 --steps 30 --warmup 5
run cfg2_bf16 --config 2 --tier bf16 --steps 30 --warmup 5
run cfg4 --config 4 --steps 30 --warmup 5
