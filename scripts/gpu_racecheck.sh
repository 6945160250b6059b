#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards), synccheck (barrier misuse) and initcheck (reads of uninitialised
# global memory) over one tiny invocation of every kernel family.  Logs: gpurun_out/{racecheck,synccheck,initcheck}.log
mkdir -p gpurun_out
timeout 120 python scripts/gpu_sanitize_small.py 2>&1 | tail -15
for tool in racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/$tool.log python scripts/gpu_sanitize_small.py 2>&1 | tail -3
  echo "$tool rc=$?"
  grep -E "SUMMARY|hazard|Barrier error|Uninitialized" gpurun_out/$tool.log | sort | uniq -c | sort -rn | head -12
done
