#!/bin/bash
# compute-sanitizer memcheck over the GAIL kernels (general program + register-tiled update), then the GPU suite and an A/B of the tiled GAIL update
set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -q -x --timeout 800 -k "gailx_shaping or gailx_depth2 or gailx_state or gail_default or gail_tuned25 or gail_mixup" > gpurun_out/r2_memcheck_gail.log 2>&1
grep -E "Invalid|ERROR SUMMARY|at .*kernel|by thread|passed|failed" gpurun_out/r2_memcheck_gail.log | head -40
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -rf > gpurun_out/r2_call5_pytest.log 2>&1
head -3 gpurun_out/r2_call5_pytest.log | cut -c1-200; tail -15 gpurun_out/r2_call5_pytest.log | cut -c1-300
for f in 0 1; do
  IL_GAIL_TILED=$f IL_TC_FUSE_L1=0 timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-eval --no-strong --no-e2e > gpurun_out/r2_gailtiled_ab_$f.json 2> gpurun_out/r2_gailtiled_ab_$f.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_gailtiled_ab_$f.json').read().strip().splitlines()[-1])
print('GAIL_TILED=$f', 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']))
PY
done
