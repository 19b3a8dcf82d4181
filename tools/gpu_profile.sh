#!/bin/bash
# round artefacts: bench lines, ncu launch list (time + DRAM bytes per launch, graphs off), one `ncu --set full` of the decode contraction
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; nvidia-smi -L >> gpurun_out/host.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -c 600 gpurun_out/bench_ours.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
SUBGC_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1300 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; head -30 gpurun_out/launches_summary.txt
SUBGC_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"h3_gemm|attention_kernel|select_reg|lstm_reduce" -s 40 -c 8 -f \
    -o gpurun_out/prof_decode python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_decode.ncu-rep --page raw --csv > gpurun_out/prof_decode.csv 2>/dev/null
python tools/ncu_pick.py gpurun_out/prof_decode.csv > gpurun_out/prof_decode_pick.txt 2>&1; cat gpurun_out/prof_decode_pick.txt | cut -c1-260
