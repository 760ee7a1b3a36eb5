#!/bin/bash
# round-2 pass o (4 GPUs): the default exchange (fused pushes) at 4 GPUs: parity + config 2 / bf16 / config 4 lines
N=${1:-4}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29541"
timeout 300 $TR tests/sharded_check.py custom > gpurun_out/r2o_w${N}_custom.log 2>&1; echo "world$N custom rc=$? $(grep -a SHARDED_OK gpurun_out/r2o_w${N}_custom.log | head -c 300)"
for t in "cfg2 --config 2" "cfg4 --config 4"; do
  tag=${t%% *}; args=${t#* }
  timeout 600 $TR bench.py --gpus $N $args --steps 30 --warmup 5 > gpurun_out/r2o_n${N}_$tag.json 2> gpurun_out/r2o_n${N}_$tag.err; echo "n$N $tag rc=$?"
  python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r2o_n${N}_$tag.json') if l.startswith('{')][0]
print('  ms', round(d['ms_per_step'],3), 'comm', d['comm']['by_kind_ms_rank0'], 'kern', d['comm']['compute_kernels_ms_per_step_max_rank'])"
done
