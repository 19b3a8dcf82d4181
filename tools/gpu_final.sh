#!/bin/bash
# final artefacts of the round: GPU test suite, smoke, both bench arms
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_ours.json"))
print({k: j[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}, j["e2e"], j["stage_ms_per_step"], {k: j["roofline"][k] for k in ("achieved", "frac", "traffic")})
PY
