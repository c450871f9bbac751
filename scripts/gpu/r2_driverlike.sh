#!/bin/bash
# what the driver runs at round end: the whole GPU suite in ONE process, smoke(), the default bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_driverlike_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_driverlike_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2_driverlike_bench.json 2> gpurun_out/r2_driverlike_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_driverlike_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('metric','value','ms_per_step','n_gpus','steps','warmup','gpu_launches','dtype','scaling')}, 'e2e', d['e2e'], 'clocks', d['clocks'], 'cpu', d['cpu_baseline'])
PY
