#!/bin/bash
# Fused first layer (producers of the tcgen05 engine compute layer 1): GPU suite with it on, then A/B bench in one call.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -40
for f in 0 1; do
  IL_TC_FUSE_L1=$f timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_fuse_ab_$f.json 2> gpurun_out/r2_fuse_ab_$f.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_fuse_ab_$f.json').read().strip().splitlines()[-1])
print('FUSE=$f', 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'dense avg ms', d['roofline'].get('avg_launch_ms'), 'frac', d['roofline'].get('frac'))
PY
done
