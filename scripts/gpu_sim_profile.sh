#!/bin/bash
mkdir -p gpurun_out
python scripts/gpu_sim_profile.py 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sim_step -c 3 -o gpurun_out/prof_sim -f python scripts/gpu_sim_profile.py > gpurun_out/ncu_sim.log 2>&1
tail -3 gpurun_out/ncu_sim.log
ls -la gpurun_out/prof_sim.ncu-rep
