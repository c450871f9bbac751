#!/bin/bash
# Usage: scripts/gpu/submit.sh <tag> <timeout_s> <script-in-repo> [gpus]
# Runs one gpurun call, retrying while the pod answers "transient / busy" (nothing charged). Log: gpurun_out/<tag>.log
tag=$1; to=$2; script=$3; gpus=${4:-1}
mkdir -p gpurun_out
for attempt in $(seq 1 40); do
  if [ "$gpus" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$to" -- bash "$script" > "gpurun_out/$tag.log" 2>&1
  else /usr/local/graft/bin/gpurun --gpus "$gpus" --timeout "$to" -- bash "$script" > "gpurun_out/$tag.log" 2>&1; fi
  rc=$?
  if grep -q "status=transient" "gpurun_out/$tag.log" || [ $rc -eq 3 ]; then echo "attempt $attempt: busy, retrying in 90 s" >> "gpurun_out/$tag.attempts"; sleep 90; continue; fi
  break
done
echo "rc=$rc attempts=$attempt" >> "gpurun_out/$tag.attempts"
