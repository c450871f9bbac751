#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_api.py -m gpu -q --timeout 600 -rf -k "adam" > gpurun_out/r2_call12_adam.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_call12_adam.log | cut -c1-300 | tail -6
timeout 600 python scripts/adam_bench.py > gpurun_out/r2_adam_variants.jsonl 2> gpurun_out/r2_adam_variants.err; cut -c1-260 gpurun_out/r2_adam_variants.jsonl; tail -3 gpurun_out/r2_adam_variants.err
timeout 600 python scripts/step_profile.py 10 > gpurun_out/r2_step_kernel_times.json 2> gpurun_out/r2_step_kernel_times.err; head -60 gpurun_out/r2_step_kernel_times.json; tail -3 gpurun_out/r2_step_kernel_times.err
for f in tests/test_gpu_loop.py tests/test_gpu_cli.py; do
  n=$(basename $f .py)
  timeout 1200 python -m pytest $f -m gpu -q --timeout 900 -rf > gpurun_out/r2_call12_$n.log 2>&1
  grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_call12_$n.log | cut -c1-300 | tail -12
done
for k in first_layer_relu_kernel gail_update_tiled_kernel adam_tma_kernel wide_tn_kernel mlp_small_forward_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r2_ncu_$k python bench.py --steps 2 --warmup 3 --start 5 --no-e2e --no-cpu-baseline --no-eval --no-strong > gpurun_out/r2_ncu_$k.log 2>&1
  tail -1 gpurun_out/r2_ncu_$k.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
