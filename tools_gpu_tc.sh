#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/pytest_gpu.log | head -30
