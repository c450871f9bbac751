#!/bin/bash
set -x
mkdir -p gpurun_out
for f in tests/test_gpu_api.py tests/test_gpu_parity.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q --timeout 600 -rf > gpurun_out/r2_call10_$n.log 2>&1
  grep -E "^FAILED|passed|failed" gpurun_out/r2_call10_$n.log | cut -c1-220 | tail -8
done
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-strong > gpurun_out/r2_bench10.json 2> gpurun_out/r2_bench10.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench10.json').read().strip().splitlines()[-1])
print('bench', 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', {k:d['roofline'].get(k) for k in ('frac','avg_launch_ms','achieved')})
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_raw.csv python bench.py --steps 2 --warmup 3 --start 5 --no-e2e --no-cpu-baseline --no-eval --no-strong > gpurun_out/r2_launches.log 2>&1
tail -2 gpurun_out/r2_launches.log | cut -c1-300
wc -l gpurun_out/r2_launches_raw.csv
python scripts/microbench.py > gpurun_out/r2_microbench.jsonl 2> gpurun_out/r2_microbench.err; cat gpurun_out/r2_microbench.jsonl | cut -c1-300; tail -2 gpurun_out/r2_microbench.err
python scripts/eval_compare.py > gpurun_out/r2_eval_compare.jsonl 2> gpurun_out/r2_eval_compare.err; cat gpurun_out/r2_eval_compare.jsonl; tail -2 gpurun_out/r2_eval_compare.err
