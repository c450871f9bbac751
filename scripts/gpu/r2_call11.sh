#!/bin/bash
set -x
mkdir -p gpurun_out
for f in tests/test_gpu_parity.py tests/test_gpu_loop.py tests/test_gpu_cli.py tests/test_gpu_api.py tests/test_gpu_tc_gemm.py; do
  n=$(basename $f .py)
  timeout 1200 python -m pytest $f -m gpu -q --timeout 900 -rf > gpurun_out/r2_call11_$n.log 2>&1
  grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_call11_$n.log | cut -c1-300 | tail -12
done
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-strong > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench11.json').read().strip().splitlines()[-1])
print('bench', 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
PY
