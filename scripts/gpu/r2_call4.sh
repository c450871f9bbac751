#!/bin/bash
# GPU suite (eval rollout graph fix, general GAIL discriminator, loop tests) + A/B of the fused first layer (FFMA2 producers) + eval / strong records
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -rf > gpurun_out/r2_call4_pytest.log 2>&1
tail -25 gpurun_out/r2_call4_pytest.log
for f in 0 1; do
  extra="--no-eval --no-strong"; [ $f = 1 ] && extra=""
  IL_TC_FUSE_L1=$f timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline $extra > gpurun_out/r2_fuse2_ab_$f.json 2> gpurun_out/r2_fuse2_ab_$f.err
  tail -3 gpurun_out/r2_fuse2_ab_$f.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_fuse2_ab_$f.json').read().strip().splitlines()[-1])
print('FUSE=$f', 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'dense avg ms', d['roofline'].get('avg_launch_ms'), 'frac', d['roofline'].get('frac'))
print('eval', d.get('eval')); print('strong', d.get('strong'))
PY
done
