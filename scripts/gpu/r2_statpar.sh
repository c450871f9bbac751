#!/bin/bash
set -x
mkdir -p gpurun_out
nproc
IL_GEMM_MODE=tf32x3 timeout 900 python scripts/statistical_parity.py 1500 64 512 > gpurun_out/r2_statistical_parity.jsonl 2> gpurun_out/r2_statistical_parity.err
cut -c1-700 gpurun_out/r2_statistical_parity.jsonl; tail -3 gpurun_out/r2_statistical_parity.err
