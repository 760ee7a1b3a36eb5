#!/bin/bash
# round-2 pass i (2 GPUs): BASELINE config 5 at FULL size (30 M nodes, 800 M edges, 3 layers, d=256, heads=4) on 2 GPUs, per-head streaming
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541"
free -g | head -2; nproc
timeout 1500 $TR bench.py --gpus 2 --config 5 --steps 3 --warmup 3 > gpurun_out/r2i_n2_cfg5_full.json 2> gpurun_out/r2i_n2_cfg5_full.err; echo "n2 cfg5 full rc=$?"
grep -a "bench +" gpurun_out/r2i_n2_cfg5_full.err | tail -8; tail -c 1500 gpurun_out/r2i_n2_cfg5_full.json; tail -5 gpurun_out/r2i_n2_cfg5_full.err | cut -c1-400
# the same configuration at 1/8 scale, streamed and not streamed, for the cost of streaming
for sh in 0 1; do
  timeout 600 $TR bench.py --gpus 2 --config 5 --scale 8 --stream-heads $sh --steps 3 --warmup 3 > gpurun_out/r2i_n2_cfg5_s8_stream$sh.json 2> gpurun_out/r2i_n2_cfg5_s8_stream$sh.err; echo "n2 cfg5/8 stream=$sh rc=$?"
  python -c "import json;d=[json.loads(l) for l in open('gpurun_out/r2i_n2_cfg5_s8_stream$sh.json') if l.startswith('{')][0];print(d['ms_per_step'], d['config'].get('hbm_peak_allocated_gb_rank0'), d['comm']['by_kind_ms_rank0'], d['breakdown_ms_per_step_rank0'])"
done
