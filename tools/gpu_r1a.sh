#!/bin/bash
# tests + bench of HEAD (fused att-phase) + A/B against the unfused step
mkdir -p gpurun_out
nproc; nvidia-smi -L | head -2
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fused.log 2>&1
tail -1 gpurun_out/bench_fused.log | cut -c1-1500
SUBGC_NO_FUSED_ATT=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_unfused.log 2>&1
tail -1 gpurun_out/bench_unfused.log | grep -o '"stage_ms_per_step": {[^}]*}'
