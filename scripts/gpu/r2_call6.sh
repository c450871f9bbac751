#!/bin/bash
# memcheck of the general GAIL program per fixture, then the GPU suite file by file (separate processes: a sticky CUDA error cannot poison later files)
set -x
mkdir -p gpurun_out
for c in gailx_shaping gailx_depth2_tanh gailx_state_only_sigmoid gail_tuned25; do
  timeout 600 compute-sanitizer --tool memcheck --print-limit 6 --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -q -x --timeout 500 -k "fixture and $c" > gpurun_out/r2_memcheck_$c.log 2>&1
  grep -E "Invalid|ERROR SUMMARY|=========     at |passed|failed|max abs err|il_b200:" gpurun_out/r2_memcheck_$c.log | head -12
done
for f in tests/test_gpu_api.py tests/test_gpu_loop.py tests/test_gpu_parity.py tests/test_gpu_tc_gemm.py tests/test_gpu_multi.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --timeout 600 -rf > gpurun_out/r2_call6_$n.log 2>&1
  head -2 gpurun_out/r2_call6_$n.log | cut -c1-160; grep -E "^FAILED|passed|failed" gpurun_out/r2_call6_$n.log | cut -c1-220 | tail -12
done
