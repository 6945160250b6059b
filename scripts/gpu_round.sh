#!/bin/bash
# One GPU visit: tests, fp64 peak, bench (ours + reference), ncu launch list + full capture of the top kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
./profiles/tools/fp64_peak | tee gpurun_out/fp64_peak.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
cp gpurun_out/fp64_peak.json profiles/fp64_peak.json
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json
tail -3 gpurun_out/bench_err.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | tee gpurun_out/bench_ref.json
if [ "$1" == "ncu" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --rollouts 8192 --no-cpu-baseline --no-mpc > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ip_solve -s 1 -c 1 -o gpurun_out/prof_ip python bench.py --steps 2 --warmup 1 --rollouts 8192 --no-cpu-baseline --no-mpc > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
fi
