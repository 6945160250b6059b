#!/bin/bash
# Quick GPU visit: all GPU tests + one bench line (no reference arm, no ncu).  Usage: gpu_quick.sh [tag]
set -x
tag=${1:-quick}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_${tag}_err.log | tee gpurun_out/bench_${tag}.json
tail -3 gpurun_out/bench_${tag}_err.log
