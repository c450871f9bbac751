#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 40 -c 5 -f -o gpurun_out/r2_ncu_final_tc_gemm python bench.py --steps 2 --warmup 3 --start 5 --no-e2e --no-cpu-baseline --no-eval --no-strong > gpurun_out/r2_ncu_final_tc_gemm.log 2>&1
tail -1 gpurun_out/r2_ncu_final_tc_gemm.log | cut -c1-200
ls -la gpurun_out/r2_ncu_final_tc_gemm.ncu-rep
