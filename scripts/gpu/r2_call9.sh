#!/bin/bash
set -x
mkdir -p gpurun_out
for f in tests/test_gpu_api.py tests/test_gpu_loop.py tests/test_gpu_parity.py tests/test_gpu_tc_gemm.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --timeout 600 -rf > gpurun_out/r2_call9_$n.log 2>&1
  grep -E "^FAILED|passed|failed" gpurun_out/r2_call9_$n.log | cut -c1-220 | tail -8
done
run() { # name env
  env $2 timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-eval --no-strong --no-e2e > gpurun_out/r2_ab9_$1.json 2> gpurun_out/r2_ab9_$1.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_ab9_$1.json').read().strip().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']))
PY
}
run all_on "IL_X=1"
run first_layer_off "IL_FIRST_LAYER_FAST=0"
run wide_tn_off "IL_WIDE_TN=0"
run all_on_again "IL_X=1"
