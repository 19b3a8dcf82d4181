#!/bin/bash
mkdir -p gpurun_out
timeout 180 python tools/gemm_check.py > gpurun_out/gemm_tc.log 2>&1; echo "rc=$?" >> gpurun_out/gemm_tc.log
SUBGC_GEMM=simt timeout 180 python tools/gemm_check.py > gpurun_out/gemm_simt.log 2>&1
cat gpurun_out/gemm_tc.log gpurun_out/gemm_simt.log
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | cut -c1-300; grep -o '"stage_ms_per_step[^}]*}' gpurun_out/bench.log; grep -o '"roofline[^}]*' gpurun_out/bench.log | cut -c1-300
