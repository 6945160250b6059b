#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_mpc.csv python bench.py --steps 1 --warmup 3 --rollouts 16384 --mpc-rollouts 16384 --no-cpu-baseline > gpurun_out/ncu_mpc.log 2>&1
tail -2 gpurun_out/ncu_mpc.log | cut -c1-300
