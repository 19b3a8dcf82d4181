#!/bin/bash
# Evidence of a round on ONE B200 (run under gpurun): GPU tests, smoke, bench lines of the three inference configurations and the
# training step, the reference arm, the ncu launch list (graphs off: every launch visible) and one `ncu --set full` capture of the
# persistent decode kernel.  Everything lands in gpurun_out/; tools/launch_summary.py / ncu_pick.py turn it into profiles/*.md.
# QUICK=1: tests, smoke, the three inference bench lines, in-kernel timeline, launch list and the ncu capture of the persistent kernel only.
R=${1:-r02}
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; nvidia-smi -L >> gpurun_out/host.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${R}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${R}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
for mode in greedy topk beam; do
  timeout 600 python bench.py --mode $mode --steps 20 --warmup 3 > gpurun_out/${R}_bench_$mode.json 2> gpurun_out/${R}_bench_$mode.err
  tail -c 300 gpurun_out/${R}_bench_$mode.json; echo
done
if [ -z "$QUICK" ]; then
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
timeout 600 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/${R}_bench_train_1gpu.json 2> gpurun_out/${R}_bench_train_1gpu.err
timeout 200 python tools/adam_bench.py > gpurun_out/${R}_adam.json 2>&1
fi
timeout 200 python tools/mega_timeline.py 128 10 > gpurun_out/${R}_mega_timeline.txt 2>&1
SUBGC_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${R}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${R}_launches.csv gpurun_out/${R}_decode_traffic.json > gpurun_out/${R}_launches_summary.txt 2>&1; head -30 gpurun_out/${R}_launches_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mega_decode_kernel -s 3 -c 1 -f -o gpurun_out/${R}_mega \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_mega.log 2>&1
ncu -i gpurun_out/${R}_mega.ncu-rep --page raw --csv > gpurun_out/${R}_mega_raw.csv 2>/dev/null
python tools/ncu_pick.py gpurun_out/${R}_mega_raw.csv > gpurun_out/${R}_mega_pick.txt 2>&1; cat gpurun_out/${R}_mega_pick.txt | cut -c1-160
[ -n "$QUICK" ] && exit 0
# front-end: one full capture of the three big encoder contractions (fusion, folded L0, one of the aggregated L1 pair): tensor-pipe activity
timeout 900 ncu --set full --clock-control none --import-source on -k regex:h3_gemm_kernel -s 36 -c 4 -f -o gpurun_out/${R}_encoder \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_encoder.log 2>&1
ncu -i gpurun_out/${R}_encoder.ncu-rep --page raw --csv > gpurun_out/${R}_encoder_raw.csv 2>/dev/null
python tools/ncu_pick.py gpurun_out/${R}_encoder_raw.csv > gpurun_out/${R}_encoder_pick.txt 2>&1; cut -c1-200 gpurun_out/${R}_encoder_pick.txt | head -60
timeout 300 python tools/graph_kernels.py greedy > gpurun_out/${R}_graph_kernels.txt 2>&1; tail -3 gpurun_out/${R}_graph_kernels.txt
timeout 300 python tools/train_profile.py 32 > gpurun_out/${R}_train_profile.txt 2>&1; head -3 gpurun_out/${R}_train_profile.txt
