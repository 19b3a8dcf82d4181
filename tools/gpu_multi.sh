#!/bin/bash
# Multi-GPU evidence on ONE box (run under `gpurun --gpus N`): the GPU tests that need two devices, sharded-inference bench, training step
# with the overlapped gradient all-reduce.  Output -> gpurun_out/<tag>_*.json
N=${1:-8}; R=${2:-r02}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_dataparallel.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29511 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_greedy_${N}gpu.json 2> gpurun_out/${R}_bench_greedy_${N}gpu.err; tail -c 400 gpurun_out/${R}_bench_greedy_${N}gpu.json; echo
run 29512 --mode topk --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_topk_${N}gpu.json 2> gpurun_out/${R}_bench_topk_${N}gpu.err; tail -c 300 gpurun_out/${R}_bench_topk_${N}gpu.json; echo
run 29513 --mode train --steps 20 --warmup 3 > gpurun_out/${R}_bench_train_${N}gpu.json 2> gpurun_out/${R}_bench_train_${N}gpu.err; tail -c 900 gpurun_out/${R}_bench_train_${N}gpu.json; echo
