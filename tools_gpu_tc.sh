#!/bin/bash
mkdir -p gpurun_out
timeout 180 python tools/gemm_check.py 2>&1 | tee gpurun_out/gemm_tc.log | grep -v mode
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|rror" gpurun_out/pytest_gpu.log | head
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | cut -c1-200; grep -o '"stage_ms_per_step[^}]*}' gpurun_out/bench.log
timeout 300 python bench.py --mode train --steps 3 --warmup 2 2>&1 | tail -1 | cut -c100-330
