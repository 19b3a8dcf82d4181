#!/bin/bash
# final artefacts: tests, smoke, both bench arms, launch list (time + DRAM bytes per launch, graphs off)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_ours.json"))
print({k: j[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}, j["e2e"], j["stage_ms_per_step"], {k: j["roofline"][k] for k in ("achieved", "frac", "traffic")})
PY
SUBGC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv | head -12
