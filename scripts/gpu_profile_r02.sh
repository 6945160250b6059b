#!/bin/bash
# Round-2 profiles (never bench values): launch list of the default bench (newton sweeps driven from the host so that
# every launch shows), one full capture of ip_solve_kernel and of newton_step_kernel, summaries for profiles/.
set -x
mkdir -p gpurun_out
CIMPC_NEWTON_HOSTLOOP=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --rollouts 8192 --mpc-rollouts 8192 --no-cpu-baseline --no-closed-loop --no-extra-configs > gpurun_out/r02_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ip_solve -s 1 -c 1 -o gpurun_out/r02_prof_ip python bench.py --steps 2 --warmup 1 --rollouts 8192 --no-cpu-baseline --no-mpc --no-extra-configs > gpurun_out/r02_ncu_ip.log 2>&1
CIMPC_NEWTON_HOSTLOOP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:newton_step -s 0 -c 2 -o gpurun_out/r02_prof_newton python bench.py --steps 1 --warmup 1 --rollouts 8192 --mpc-rollouts 4096 --no-cpu-baseline --no-closed-loop --no-extra-configs > gpurun_out/r02_ncu_newton.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02_prof_ip.ncu-rep --source ip_solve > gpurun_out/r02_ip_summary.txt 2>&1
python scripts/ncu_summary.py gpurun_out/r02_prof_newton.ncu-rep > gpurun_out/r02_newton_summary.txt 2>&1
ls -la gpurun_out | tail -12
