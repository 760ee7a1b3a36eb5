#!/bin/bash
# round-2 pass e (1 GPU): dW accuracy A/B (cross-term tile, promotion group), stream tests, per-rank kernel times by emulation
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for lib in plotpointe-gat-recommendation_b200/libb200gat.so ab/libb200gat_noxt.so ab/libb200gat_xt_g1.so; do
  tag=$(basename $lib .so)
  B200GAT_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gemm_tc.py -q -k "config1_shape or against_reference_golden or gemm" > gpurun_out/r2e_dw_$tag.log 2>&1
  echo "dw $tag: $(tail -1 gpurun_out/r2e_dw_$tag.log)"
  cp gpurun_out/parity_report.json gpurun_out/r2e_parity_$tag.json
  B200GAT_LIB=$PWD/$lib timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-next-rows > gpurun_out/r2e_cfg2_$tag.json 2>gpurun_out/r2e_cfg2_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/r2e_cfg2_$tag.json'));print(d['ms_per_step'], d['breakdown_ms_per_step'])"
done
timeout 600 python -m pytest tests/test_gpu_stream.py -q > gpurun_out/r2e_stream.log 2>&1; echo "stream: $(tail -1 gpurun_out/r2e_stream.log)"
timeout 900 python tools/diag/rank_emulate.py 1 2 4 8 > gpurun_out/r2e_emulate_f32.log 2>&1; cat gpurun_out/r2e_emulate_f32.log | tail -8
timeout 900 python tools/diag/rank_emulate.py 1 8 --tier bf16 > gpurun_out/r2e_emulate_bf16.log 2>&1; cat gpurun_out/r2e_emulate_bf16.log | tail -4
