#!/bin/bash
# round-2 pass f (1 GPU): which tensor-core GEMM carries the config-1 gradient error (B200GAT_TC_PARTS), stream tests, emulation
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for parts in 31 30 29 27 23 15 1 2 4; do
  B200GAT_TC_PARTS=$parts timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "config1_shape and custom" > gpurun_out/r2f_parts_$parts.log 2>&1
  cp gpurun_out/parity_report.json gpurun_out/r2f_parity_parts_$parts.json
  echo "parts $parts: $(tail -1 gpurun_out/r2f_parts_$parts.log)"
done
timeout 600 python -m pytest tests/test_gpu_stream.py -q > gpurun_out/r2f_stream.log 2>&1; echo "stream: $(tail -1 gpurun_out/r2f_stream.log)"
timeout 900 python tools/diag/rank_emulate.py 1 2 4 8 > gpurun_out/r2f_emulate_f32.log 2>&1; tail -4 gpurun_out/r2f_emulate_f32.log | cut -c1-1500
timeout 900 python tools/diag/rank_emulate.py 1 8 --tier bf16 > gpurun_out/r2f_emulate_bf16.log 2>&1; tail -2 gpurun_out/r2f_emulate_bf16.log | cut -c1-1500
