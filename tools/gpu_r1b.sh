#!/bin/bash
# tests + bench + ncu launch list (graphs off so that every launch is visible)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | grep -o '"value": [0-9.]*\|"stage_ms_per_step": {[^}]*}\|"frac": [0-9.]*' | head -4
if [ -n "$AB" ]; then
SUBGC_H3=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"stage_ms_per_step": {[^}]*}'
fi
SUBGC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv | head -60
