#!/bin/bash
set -x
mkdir -p gpurun_out
for f in tests/test_gpu_api.py tests/test_gpu_loop.py tests/test_gpu_parity.py tests/test_gpu_tc_gemm.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --timeout 600 -rf > gpurun_out/r2_call8_$n.log 2>&1
  grep -E "^FAILED|passed|failed" gpurun_out/r2_call8_$n.log | cut -c1-220 | tail -8
done
for f in 0 1; do
  extra="--no-eval --no-strong --no-e2e"; [ $f = 1 ] && extra="--no-strong"
  IL_HEAD_FUSED=$f timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline $extra > gpurun_out/r2_headfused_ab_$f.json 2> gpurun_out/r2_headfused_ab_$f.err
  tail -2 gpurun_out/r2_headfused_ab_$f.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_headfused_ab_$f.json').read().strip().splitlines()[-1])
print('HEAD_FUSED=$f', 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'eval', d.get('eval'))
PY
done
