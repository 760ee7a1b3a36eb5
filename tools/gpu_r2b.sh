#!/bin/bash
# round-2 second GPU pass (2 GPUs): sharded parity over the peer fabric, N=2 bench lines, dW promotion A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541"
for k in pyg custom; do
  timeout 300 python tests/sharded_check.py $k > gpurun_out/r2b_w1_$k.log 2>&1; echo "world1 $k rc=$? $(tail -1 gpurun_out/r2b_w1_$k.log | head -c 200)"
  timeout 300 $TR tests/sharded_check.py $k > gpurun_out/r2b_w2_$k.log 2>&1; echo "world2 $k rc=$? $(grep -a SHARDED_OK gpurun_out/r2b_w2_$k.log | head -c 200)"
done
for lib in plotpointe-gat-recommendation_b200/libb200gat.so ab/libb200gat_dwg8.so ab/libb200gat_dwg1.so; do
  tag=$(basename $lib .so)
  B200GAT_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "config1_shape or against_reference_golden" > gpurun_out/r2b_dw_$tag.log 2>&1
  echo "dw $tag: $(tail -1 gpurun_out/r2b_dw_$tag.log)"
  cp gpurun_out/parity_report.json gpurun_out/r2b_parity_$tag.json
  B200GAT_LIB=$PWD/$lib timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-next-rows > gpurun_out/r2b_cfg2_$tag.json 2>gpurun_out/r2b_cfg2_$tag.err
done
timeout 600 $TR bench.py --gpus 2 --config 2 --steps 20 --warmup 5 > gpurun_out/r2b_n2_cfg2.json 2> gpurun_out/r2b_n2_cfg2.err; echo "n2 cfg2 rc=$?"
timeout 600 $TR bench.py --gpus 2 --config 4 --steps 20 --warmup 5 > gpurun_out/r2b_n2_cfg4.json 2> gpurun_out/r2b_n2_cfg4.err; echo "n2 cfg4 rc=$?"
timeout 600 $TR bench.py --gpus 2 --config 2 --tier bf16 --steps 20 --warmup 5 > gpurun_out/r2b_n2_cfg2_bf16.json 2> gpurun_out/r2b_n2_cfg2_bf16.err; echo "n2 cfg2 bf16 rc=$?"
tail -c 1500 gpurun_out/r2b_n2_cfg2.json
