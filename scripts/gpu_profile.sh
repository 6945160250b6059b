#!/bin/bash
# ncu launch list + one full capture of ip_solve_kernel (never a bench value)
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --rollouts 8192 --no-cpu-baseline --no-mpc > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ip_solve -s 1 -c 1 -o gpurun_out/prof_ip python bench.py --steps 2 --warmup 1 --rollouts 8192 --no-cpu-baseline --no-mpc > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
