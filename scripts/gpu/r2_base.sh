#!/bin/bash
# Round-2 baseline on this round's boxes: GPU suite, bench line, ncu --set full of the GAIL kernels (no capture existed in round 1).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_base_bench.json 2> gpurun_out/r2_base_bench.err
tail -c 700 gpurun_out/r2_base_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gail_update|gail_reward' -s 4 -c 2 -o gpurun_out/r2_gail_base python bench.py --steps 2 --warmup 3 --start 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu_gail.log 2>&1
tail -2 gpurun_out/r2_ncu_gail.log
