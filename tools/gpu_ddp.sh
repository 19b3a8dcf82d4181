#!/bin/bash
# 2-GPU checks: gradient all-reduce parity, sharded-inference bench (weak scaling), training-step bench
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_2gpu.json
cut -c1-330 gpurun_out/bench_2gpu.json
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
timeout 300 python bench.py --mode beam --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"stage_ms_per_step": {[^}]*}' | head -3
timeout 300 python bench.py --mode topk --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"stage_ms_per_step": {[^}]*}' | head -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --mode train --steps 3 --warmup 2 2>&1 | tail -1 | cut -c1-400
