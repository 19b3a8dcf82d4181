#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_tc.json 2> gpurun_out/bench_err.log
tail -1 gpurun_out/bench_r01_tc.json | cut -c1-250
SUBGC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_tc.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
SUBGC_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"umma_gemm|attention_kernel|select_kernel|lstm_reduce" -s 60 -c 8 -f -o gpurun_out/prof_tc \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
