#!/bin/bash
# compute-sanitizer over the GPU parity tests: memcheck on the fixture tests, racecheck + synccheck on the tcgen05 / mbarrier engine tests
set -x
mkdir -p gpurun_out
export PYTEST_ADDOPTS="-p no:cacheprovider"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 1400 -k "fixture" > gpurun_out/r2_memcheck_parity.log 2>&1; echo "memcheck parity rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_memcheck_parity.log | tail -4
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_tc_gemm.py -m gpu -q --timeout 1100 > gpurun_out/r2_memcheck_tc_gemm.log 2>&1; echo "memcheck tc_gemm rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_memcheck_tc_gemm.log | tail -4
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_tc_gemm.py -m gpu -q --timeout 1400 -k "matches_fp64 or golden_cases" > gpurun_out/r2_racecheck_tc_gemm.log 2>&1; echo "racecheck tc_gemm rc=$?"
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/r2_racecheck_tc_gemm.log | tail -4
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_tc_gemm.py -m gpu -q --timeout 800 -k "golden_cases" > gpurun_out/r2_synccheck_tc_gemm.log 2>&1; echo "synccheck tc_gemm rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_synccheck_tc_gemm.log | tail -4
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_api.py -m gpu -q --timeout 800 -k "adam" > gpurun_out/r2_racecheck_adam.log 2>&1; echo "racecheck adam rc=$?"
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/r2_racecheck_adam.log | tail -4
