#!/bin/bash
# final verification of the round: whole GPU suite (one process per file), smoke(), the full bench line, the ncu launch list with DRAM counters, in-step profile
set -x
mkdir -p gpurun_out
for f in tests/test_gpu_parity.py tests/test_gpu_tc_gemm.py tests/test_gpu_api.py tests/test_gpu_loop.py tests/test_gpu_cli.py tests/test_gpu_multi.py; do
  n=$(basename $f .py)
  timeout 1200 python -m pytest $f -m gpu -q --timeout 900 -rf > gpurun_out/r2_call23_$n.log 2>&1
  grep -E "^FAILED|^ERROR|passed|failed|skipped" gpurun_out/r2_call23_$n.log | cut -c1-300 | tail -12
done
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench_final2.json 2> gpurun_out/r2_bench_final2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_final2.json').read().strip().splitlines()[-1])
print('final', 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', {k:d['roofline'].get(k) for k in ('frac','avg_launch_ms','achieved','traffic','algorithmic_bytes_per_launch')}, 'cpu', d.get('cpu_baseline'), 'eval', d['eval']['value'], 'launches', d['gpu_launches'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_final2_raw.csv python bench.py --steps 2 --warmup 3 --start 5 --no-e2e --no-cpu-baseline --no-eval --no-strong > gpurun_out/r2_launches_final2.log 2>&1
wc -l gpurun_out/r2_launches_final2_raw.csv
timeout 600 python scripts/step_profile.py 10 > gpurun_out/r2_step_kernel_times.json 2> gpurun_out/r2_step_kernel_times.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_step_kernel_times.json'))
print('step ms', d['ms_per_step_without_profiler'], 'kernel ms', d['kernel_ms_per_step'])
for k in d['kernels'][:12]: print(f"{k['kernel'][:48]:48s} x{k['launches_per_step']:5.1f} {k['us_per_step']:8.1f} us {100*k['share']:5.1f}%")
PY
