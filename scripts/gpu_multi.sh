#!/bin/bash
# 2-GPU check of the torchrun path (bench only)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>gpurun_out/bench2_err.log | tee gpurun_out/bench_2gpu.json
tail -5 gpurun_out/bench2_err.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 | tee gpurun_out/bench_2gpu_ref.json
