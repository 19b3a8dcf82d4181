#!/bin/bash
# compute-sanitizer over smoke() (full model dimensions: tcgen05 contractions, PDL-chained per-stage kernels, the persistent decode kernel,
# the whole call replayed as a CUDA graph).  Logs -> gpurun_out/sanitize_*.log; summarised under profiles/.
mkdir -p gpurun_out
export SUBGC_MEGA_TIMEOUT_S=600   # the persistent kernel is ~100x slower under the tools; its 4 s abort guard would fire
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|hazard" gpurun_out/sanitize_$tool.log | head -8
done
# the same with the whole-step graph and the persistent kernel switched off (per-stage kernels under PDL)
SUBGC_MEGA=0 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_racecheck_perstage.log 2>&1
echo "== racecheck per-stage rc=$?"; grep -E "RACECHECK SUMMARY|smoke ok|hazard" gpurun_out/sanitize_racecheck_perstage.log | head -5
