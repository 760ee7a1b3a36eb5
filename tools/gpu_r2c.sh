#!/bin/bash
# 1-GPU pass: streaming test, dW promotion A/B, cfg5 at 1/64 scale
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stream.py tests/test_gpu_gemm_tc.py -q -x > gpurun_out/r2c_stream.log 2>&1; echo "stream+gemm tests: $(tail -1 gpurun_out/r2c_stream.log)"
cp gpurun_out/parity_report.json gpurun_out/r2c_parity_stream.json 2>/dev/null
for k in pyg custom; do
  timeout 300 python tests/sharded_check.py $k > gpurun_out/r2c_w1_$k.log 2>&1; echo "world1 $k rc=$? $(tail -1 gpurun_out/r2c_w1_$k.log | head -c 300)"
done
for lib in plotpointe-gat-recommendation_b200/libb200gat.so ab/libb200gat_dwg8.so ab/libb200gat_dwg1.so; do
  tag=$(basename $lib .so)
  B200GAT_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "config1_shape or against_reference_golden" > gpurun_out/r2c_dw_$tag.log 2>&1
  echo "dw $tag: $(tail -1 gpurun_out/r2c_dw_$tag.log)"
  cp gpurun_out/parity_report.json gpurun_out/r2c_parity_$tag.json
  B200GAT_LIB=$PWD/$lib timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-next-rows > gpurun_out/r2c_cfg2_$tag.json 2>gpurun_out/r2c_cfg2_$tag.err
done
timeout 900 python bench.py --config 5 --scale 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_cfg5_s64.json 2> gpurun_out/r2c_cfg5_s64.err; echo "cfg5/64 rc=$?"
tail -c 1200 gpurun_out/r2c_cfg5_s64.json; tail -5 gpurun_out/r2c_cfg5_s64.err
