#!/bin/bash
# round-2 pass d (2 GPUs): full GPU suite on the rebuilt library, sharded parity (world 1 and 2), N=2 bench lines, cfg5 at 1/64 scale
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2d_pytest.log)"
cp gpurun_out/parity_report.json gpurun_out/r2d_parity_report.json 2>/dev/null
for k in pyg custom; do
  timeout 300 python tests/sharded_check.py $k > gpurun_out/r2d_w1_$k.log 2>&1; echo "world1 $k rc=$? $(tail -1 gpurun_out/r2d_w1_$k.log | head -c 300)"
  timeout 300 $TR tests/sharded_check.py $k > gpurun_out/r2d_w2_$k.log 2>&1; echo "world2 $k rc=$? $(grep -a SHARDED_OK gpurun_out/r2d_w2_$k.log | head -c 300)"
done
timeout 300 python bench.py --config 2 --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows > gpurun_out/r2d_cfg2.json 2>gpurun_out/r2d_cfg2.err; echo "cfg2 rc=$?"
timeout 600 $TR bench.py --gpus 2 --config 2 --steps 20 --warmup 5 > gpurun_out/r2d_n2_cfg2.json 2> gpurun_out/r2d_n2_cfg2.err; echo "n2 cfg2 rc=$?"
timeout 600 $TR bench.py --gpus 2 --config 4 --steps 20 --warmup 5 > gpurun_out/r2d_n2_cfg4.json 2> gpurun_out/r2d_n2_cfg4.err; echo "n2 cfg4 rc=$?"
timeout 600 $TR bench.py --gpus 2 --config 2 --tier bf16 --steps 20 --warmup 5 > gpurun_out/r2d_n2_cfg2_bf16.json 2> gpurun_out/r2d_n2_cfg2_bf16.err; echo "n2 cfg2 bf16 rc=$?"
timeout 900 python bench.py --config 5 --scale 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_cfg5_s64.json 2> gpurun_out/r2d_cfg5_s64.err; echo "cfg5/64 rc=$?"
timeout 900 $TR bench.py --gpus 2 --config 5 --scale 64 --steps 5 --warmup 3 > gpurun_out/r2d_n2_cfg5_s64.json 2> gpurun_out/r2d_n2_cfg5_s64.err; echo "n2 cfg5/64 rc=$?"
for f in r2d_cfg2 r2d_n2_cfg2 r2d_n2_cfg4 r2d_n2_cfg2_bf16 r2d_cfg5_s64 r2d_n2_cfg5_s64; do echo "== $f"; head -c 400 gpurun_out/$f.json; echo; tail -3 gpurun_out/$f.err; done
