#!/bin/bash
for m in 0 1 2 16 32 64 85; do
  r=$(SUBGC_SKIP=$m timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -o '"decode": [0-9][0-9.]*' | tr '\n' ' ')
  echo "fused skip=$m $r"
done
