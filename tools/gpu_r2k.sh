#!/bin/bash
# round-2 pass k (2 GPUs): BASELINE config 5 at FULL size (30 M nodes, 800 M edges, 3 layers, d=256, heads=4) on 2 GPUs, per-head streaming
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541"
timeout 1500 $TR bench.py --gpus 2 --config 5 --steps 3 --warmup 3 > gpurun_out/r2k_n2_cfg5_full.json 2> gpurun_out/r2k_n2_cfg5_full.err; echo "n2 cfg5 full rc=$?"
grep -a "bench +" gpurun_out/r2k_n2_cfg5_full.err | tail -8; tail -c 2500 gpurun_out/r2k_n2_cfg5_full.json; grep -a "Error\|error" gpurun_out/r2k_n2_cfg5_full.err | head -5 | cut -c1-400
