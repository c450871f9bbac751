#!/bin/bash
# 2-GPU call: the library's own NCCL communicator + fused return all-reduce test, the bench with strong / eval records at N = 2
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 500 -rf > gpurun_out/r2_multi2_test.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed|skipped" gpurun_out/r2_multi2_test.log | cut -c1-300 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -1 gpurun_out/r2_bench_2gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=2', 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'strong', d.get('strong'), 'eval', {k:d['eval'].get(k) for k in ('value','return_allreduce_us','episodes')})"
tail -3 gpurun_out/r2_bench_2gpu.err
