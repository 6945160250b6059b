#!/bin/bash
# ncu full capture (with source) of the Monte-Carlo drop-box simulator step: its duration is the straggler tail
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sim_step -s 2 -c 1 -f -o gpurun_out/prof_sim_box python scripts/gpu_sim_profile.py > gpurun_out/ncu_sim_box.log 2>&1
tail -3 gpurun_out/ncu_sim_box.log | cut -c1-200
ls -la gpurun_out/prof_sim_box.ncu-rep
