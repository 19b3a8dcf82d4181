#!/bin/bash
# one `ncu --set full` capture of the kernels matching $1 (regex), skipping $2 launches, $3 captures; graphs off
mkdir -p gpurun_out
K=${1:-att_phase}; S=${2:-20}; C=${3:-2}
SUBGC_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -f -o gpurun_out/prof_$K \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$K.log 2>&1
ncu -i gpurun_out/prof_$K.ncu-rep --page raw --csv > gpurun_out/prof_$K.csv 2>/dev/null
python tools/ncu_pick.py gpurun_out/prof_$K.csv
