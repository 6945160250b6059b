#!/bin/bash
# ncu full capture of the first two newton_step_kernel launches of the MPC-step leg (INIT: every rollout solves its KKT system)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:newton_step -s 0 -c 2 -o gpurun_out/prof_newton -f python bench.py --steps 1 --warmup 1 --rollouts 4096 --mpc-rollouts 16384 --no-cpu-baseline --no-closed-loop > gpurun_out/ncu_newton.log 2>&1
tail -2 gpurun_out/ncu_newton.log | cut -c1-200
ls -la gpurun_out/prof_newton.ncu-rep
