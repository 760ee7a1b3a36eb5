#!/bin/bash
# 1-GPU pass: world-1 sharded path over the fabric kernels, bf16 GEMM tests, dW promotion A/B, config 3 with the fused backward
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for k in pyg custom; do
  timeout 300 python tests/sharded_check.py $k > gpurun_out/r2b_w1_$k.log 2>&1; echo "world1 $k rc=$? $(tail -1 gpurun_out/r2b_w1_$k.log | head -c 300)"
done
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -x > gpurun_out/r2b_gemm.log 2>&1; echo "gemm tests: $(tail -1 gpurun_out/r2b_gemm.log)"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "bf16 or c256" > gpurun_out/r2b_bf16.log 2>&1; echo "bf16 tier tests: $(tail -1 gpurun_out/r2b_bf16.log)"
for lib in "" ab/libb200gat_dwg8.so ab/libb200gat_dwg1.so; do
  tag=$(basename ${lib:-default})
  B200GAT_LIB=${lib:+$PWD/$lib} timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "config1_shape or against_reference_golden" > gpurun_out/r2b_dw_$tag.log 2>&1
  echo "dw $tag: $(tail -1 gpurun_out/r2b_dw_$tag.log)"
  cp gpurun_out/parity_report.json gpurun_out/r2b_parity_$tag.json
  B200GAT_LIB=${lib:+$PWD/$lib} timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-next-rows > gpurun_out/r2b_cfg2_$tag.json 2>gpurun_out/r2b_cfg2_$tag.err
done
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_cfg3.json 2> gpurun_out/r2b_cfg3.err; echo "cfg3 rc=$?"
tail -c 800 gpurun_out/r2b_cfg3.json
