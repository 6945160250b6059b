#!/bin/bash
# End-of-round evidence: per-config throughput, bench (ours + reference arm), ncu launch list and full capture.
set -x
mkdir -p gpurun_out
python scripts/gpu_configs.py 2>/dev/null | tee gpurun_out/configs.jsonl
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | tee gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --rollouts 8192 --no-cpu-baseline --no-mpc > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ip_solve -s 1 -c 1 -f -o gpurun_out/prof_ip python bench.py --steps 2 --warmup 1 --rollouts 8192 --no-cpu-baseline --no-mpc > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
