#!/bin/bash
# 8-GPU call: BASELINE configs 4-5 (SAC ant 4096 / 8 GPUs, GAIL walker2d 8192 / 8 GPUs + return all-reduce) and the bench line with strong / eval records at N = 8
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/run_configs.py > gpurun_out/r2_configs_8gpu.jsonl 2> gpurun_out/r2_configs_8gpu.err
cut -c1-300 gpurun_out/r2_configs_8gpu.jsonl; tail -3 gpurun_out/r2_configs_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
tail -1 gpurun_out/r2_bench_8gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=8', 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'strong', d.get('strong'), 'eval', {k:d['eval'].get(k) for k in ('value','return_allreduce_us','episodes')})"
tail -3 gpurun_out/r2_bench_8gpu.err
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 280 -rf 2>&1 | tail -3
