#!/bin/bash
set -x
mkdir -p gpurun_out
for f in tests/test_gpu_tc_gemm.py tests/test_gpu_parity.py tests/test_gpu_loop.py tests/test_gpu_api.py; do
  n=$(basename $f .py)
  timeout 1200 python -m pytest $f -m gpu -q --timeout 900 -rf > gpurun_out/r2_call21_$n.log 2>&1
  grep -E "^FAILED|^ERROR|passed|failed|skipped" gpurun_out/r2_call21_$n.log | cut -c1-300 | tail -12
done
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-strong --no-eval > gpurun_out/r2_ab21_$tag.json 2> gpurun_out/r2_ab21_$tag.err
  python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/r2_ab21_$tag.json').read().strip().splitlines()[-1])
  print('$tag', 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'dense avg ms', round(d['roofline']['avg_launch_ms'],4))
except Exception as e:
  print('$tag', 'FAILED', e); print(open('gpurun_out/r2_ab21_$tag.err').read()[-1500:])
PY
}
run a IL_DUMMY=1
run b IL_DUMMY=1
timeout 600 python scripts/step_profile.py 10 > gpurun_out/r2_step_kernel_times.json 2> gpurun_out/r2_step_kernel_times.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_step_kernel_times.json'))
print('step ms', d['ms_per_step_without_profiler'], 'kernel ms', d['kernel_ms_per_step'])
for k in d['kernels'][:14]: print(f"{k['kernel'][:48]:48s} x{k['launches_per_step']:5.1f} {k['us_per_step']:8.1f} us {100*k['share']:5.1f}%")
PY
for k in wide_tn_kernel mlp_small_forward_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -f -o gpurun_out/r2_ncu2_$k python bench.py --steps 2 --warmup 3 --start 5 --no-e2e --no-cpu-baseline --no-eval --no-strong > gpurun_out/r2_ncu2_$k.log 2>&1
  tail -1 gpurun_out/r2_ncu2_$k.log | cut -c1-200
done
