#!/bin/bash
mkdir -p gpurun_out
timeout 180 python tools/gemm_check.py 2>&1 | grep "N=4000 K=4000\|N=1024\|N=9488"
SUBGC_TC_NOCOLLECT=1 timeout 180 python tools/gemm_check.py 2>&1 | grep "N=4000 K=4000\|N=1024\|N=9488"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|rror" gpurun_out/pytest_gpu.log | head
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
grep -o '"value": [0-9.]*' gpurun_out/bench.log | head -3; grep -o '"stage_ms_per_step[^}]*}' gpurun_out/bench.log; grep -o '"e2e": {[^}]*}' gpurun_out/bench.log
SUBGC_TC_NOCOLLECT=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"stage_ms_per_step[^}]*}'
