#!/bin/bash
# round-2 pass ai (1 GPU): final records -- driver sequence (smoke, GPU suite, default bench line, reference arm), then an
# ncu --set full capture of the two edge kernels of the final code (DRAM traffic per launch for profiles/traffic.json)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
bash tools/gpu_final.sh
B="python bench.py --config 2 --steps 2 --warmup 3 --no-cpu-baseline --no-next-rows"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_ -c 4 -f -o gpurun_out/r2ai_edge $B > gpurun_out/r2ai_ncu_edge.log 2>&1
echo "edge capture rc=$?"
[ -f gpurun_out/r2ai_edge.ncu-rep ] && ncu -i gpurun_out/r2ai_edge.ncu-rep --page raw --csv > gpurun_out/r2ai_edge_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2ai_edge_raw.csv "r2ai: edge kernels of the final code, config 2" | tail -6
