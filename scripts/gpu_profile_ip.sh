#!/bin/bash
# ncu --set full of ONE launch of ip_solve_kernel (81 920 quadruped subproblems) + the launch list of the bench (host-driven
# Newton sweeps so that every launch shows).  Never bench values.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ip_solve -c 1 -f -o gpurun_out/r02f_prof_ip python scripts/gpu_ip_ab_profile.py quadruped 81920 v2 > gpurun_out/r02f_ncu_ip.log 2>&1
tail -2 gpurun_out/r02f_ncu_ip.log
CIMPC_NEWTON_HOSTLOOP=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 1 --rollouts 8192 --mpc-rollouts 8192 --no-cpu-baseline --no-closed-loop --no-extra-configs > gpurun_out/r02f_ncu_launches.log 2>&1
tail -2 gpurun_out/r02f_ncu_launches.log
ls -la gpurun_out/r02f*
