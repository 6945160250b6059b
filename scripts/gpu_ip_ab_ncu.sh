#!/bin/bash
# ncu --set full of one launch of ip_solve_kernel (v2) and ip_solve2_kernel (v3) on the same 81 920-subproblem batch
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ip_solve -c 2 -f -o gpurun_out/prof_ip_ab python scripts/gpu_ip_ab_profile.py quadruped 81920 > gpurun_out/ncu_ip_ab.log 2>&1
tail -3 gpurun_out/ncu_ip_ab.log
ls -la gpurun_out/*.ncu-rep
