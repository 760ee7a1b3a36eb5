#!/bin/bash
# round-2 pass am (1 GPU): cp.async rings in the warp-per-row edge kernels (heads 4): config-3 step A/B, parity tests with the ring build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for lib in ab/lib_e_g0.so ab/lib_e_g1.so ab/lib_e_g2.so ab/lib_e_g3.so; do
  echo "== $lib"
  B200GAT_LIB=$PWD/$lib CASES="h4 bf16 bpr" timeout 200 python tools/diag/config3_timing.py 2>&1 | tail -1 | cut -c1-420
done > gpurun_out/r2am_cfg3_ring_ab.log 2>&1
cat gpurun_out/r2am_cfg3_ring_ab.log
B200GAT_LIB=$PWD/ab/lib_e_g1.so timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stream.py -x -q > gpurun_out/r2am_tests_g1.log 2>&1; echo "g1 tests rc=$? $(tail -1 gpurun_out/r2am_tests_g1.log)"
