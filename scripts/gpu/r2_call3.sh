#!/bin/bash
# GPU suite (new: eval rollout graph, prefill / PWIL relabel loops, bench-config loop, tuned GAIL fixtures) + ncu --set full of the fused launches
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -c 5 -o gpurun_out/r2_fuse_prof python bench.py --steps 2 --warmup 3 --start 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_fuse_prof.log 2>&1
tail -3 gpurun_out/r2_fuse_prof.log
ls -la gpurun_out/
