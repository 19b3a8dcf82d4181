#!/bin/bash
# tests + bench (+ optional env A/B given as "VAR=1" arguments, one bench line each)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -2 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | grep -o '"value": [0-9.]*\|"stage_ms_per_step": {[^}]*}\|"frac": [0-9.]*' | head -4
for v in "$@"; do
  echo "--- $v"
  env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"stage_ms_per_step": {[^}]*}'
done
