#!/bin/bash
# ncu --set full of the first two newton_step launches of one solve (launch 1: every rollout solves its KKT system).
mkdir -p gpurun_out
K=${1:-newton_step_cta}
CIMPC_NEWTON_HOSTLOOP=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:$K -c 2 -f -o gpurun_out/newton_step_full \
  python scripts/gpu_mpc_solve.py --solves 0 --rollouts ${2:-16384} > gpurun_out/ncu_newton_full.log 2>&1
tail -3 gpurun_out/ncu_newton_full.log | cut -c1-200
ls -la gpurun_out/newton_step_full.ncu-rep
