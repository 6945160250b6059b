#!/bin/bash
# compute-sanitizer memcheck over the C-ABI life-cycle tests and one small case of every kernel family
# (SURVEY §5).  Slow (10-50× under the tool): small batches only.  Log: gpurun_out/memcheck.log
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log \
  python -m pytest -x -q -m gpu tests/test_gpu_abi.py "tests/test_gpu_linearize.py::test_solves_after_device_linearization_match_uploaded" \
  "tests/test_gpu_newton.py" -k "abi or linearization or quadruped" 2>&1 | tail -5
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/memcheck.log | head -20
