#!/bin/bash
# round-2 first GPU pass: parity suite, bench lines of configs 1-4, CPU arms
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt; nproc >> gpurun_out/r2a_gpu.txt; free -g >> gpurun_out/r2a_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
for c in 2 1 4; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 $( [ $c != 2 ] && echo --no-cpu-baseline ) > gpurun_out/r2a_cfg$c.json 2> gpurun_out/r2a_cfg$c.err; echo "cfg$c rc=$?"
done
for l in bpr bce; do
  timeout 600 python bench.py --config 3 --loss $l --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_cfg3_$l.json 2> gpurun_out/r2a_cfg3_$l.err; echo "cfg3 $l rc=$?"
done
timeout 900 python bench.py --impl reference --config 2 --steps 2 --warmup 1 > gpurun_out/r2a_ref_cfg2.json 2> gpurun_out/r2a_ref_cfg2.err; echo "ref2 rc=$?"
timeout 300 python bench.py --impl reference --config 1 --steps 2 --warmup 1 > gpurun_out/r2a_ref_cfg1.json 2> gpurun_out/r2a_ref_cfg1.err; echo "ref1 rc=$?"
head -c 600 gpurun_out/r2a_cfg2.json
