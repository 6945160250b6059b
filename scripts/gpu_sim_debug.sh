#!/bin/bash
python scripts/gpu_sim_debug.py 2>&1 | tail -5
python -m pytest tests/test_gpu_simulator.py -m gpu -q -x -k flamingo 2>&1 | grep -E "assert|Error|same|worst|^E" | head -12
