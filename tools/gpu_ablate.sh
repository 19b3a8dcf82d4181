#!/bin/bash
# in-graph cost of each decode-step kernel: decode stage time with that kernel not launched (results are wrong, timing only)
for m in 0 1 2 16 32 64 128 255; do
  r=$(SUBGC_SKIP=$m $EXTRA timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"decode": [0-9][0-9.]*' | tr '\n' ' ')
  echo "skip=$m $r"
done
