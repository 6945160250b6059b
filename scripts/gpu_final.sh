#!/bin/bash
# End-of-round evidence: full GPU suite, bench (ours + reference arm), final ncu capture + launch list (never bench values).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_final_err.log > gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final_err.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json
bash scripts/gpu_profile_ip.sh > /dev/null 2>&1
ls -la gpurun_out | tail -8
