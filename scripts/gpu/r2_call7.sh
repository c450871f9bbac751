#!/bin/bash
set -x
mkdir -p gpurun_out
IL_DEBUG_SYNC=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 500 -k "fixture and gailx" > gpurun_out/r2_call7_gailx_dbg.log 2>&1
grep -E "il_b200:|passed|failed|max abs err" gpurun_out/r2_call7_gailx_dbg.log | cut -c1-400 | head -20
for f in tests/test_gpu_api.py tests/test_gpu_loop.py tests/test_gpu_parity.py tests/test_gpu_tc_gemm.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --timeout 600 -rf > gpurun_out/r2_call7_$n.log 2>&1
  head -2 gpurun_out/r2_call7_$n.log | cut -c1-160; grep -E "^FAILED|passed|failed" gpurun_out/r2_call7_$n.log | cut -c1-220 | tail -12
done
for f in 0 1; do
  IL_ADAM_TMA=$f timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-eval --no-strong --no-e2e > gpurun_out/r2_adamtma_ab_$f.json 2> gpurun_out/r2_adamtma_ab_$f.err
  tail -2 gpurun_out/r2_adamtma_ab_$f.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_adamtma_ab_$f.json').read().strip().splitlines()[-1])
print('ADAM_TMA=$f', 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']))
PY
done
