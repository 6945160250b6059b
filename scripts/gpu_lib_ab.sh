#!/bin/bash
# A/B of two builds of the library on the per-config IP throughput, the MPC step, and a bitwise comparison of solver outputs.
# usage: gpu_lib_ab.sh OTHER_LIB.so   (the default build is lib/libcimpc_b200.so)
mkdir -p gpurun_out
O=$1
fmt='import sys,json; [print(d["robot"],d["mode"],d["subproblems"],round(d["ms"],3),"ms",round(d["subproblems_per_s"]/1e6,2),"M/s") for d in map(json.loads,sys.stdin)]'
for rep in 1 2; do
echo "== other ($O)"; CIMPC_B200_LIB=$O python scripts/gpu_configs.py 2>/dev/null | python -c "$fmt"
echo "== default"; python scripts/gpu_configs.py 2>/dev/null | python -c "$fmt"
done
echo "== mpc other"; CIMPC_B200_LIB=$O python scripts/gpu_mpc_solve.py --solves 5 | tail -2
echo "== mpc default"; python scripts/gpu_mpc_solve.py --solves 5 | tail -2
python scripts/gpu_ip_dump.py gpurun_out/dump_a.npz > /dev/null
CIMPC_B200_LIB=$O python scripts/gpu_ip_dump.py gpurun_out/dump_b.npz > /dev/null
python - <<'PY'
import numpy as np
a, b = np.load("gpurun_out/dump_a.npz"), np.load("gpurun_out/dump_b.npz")
bad = [k for k in a.files if not (np.array_equal(a[k], b[k], equal_nan=True) if a[k].dtype.kind == "f" else np.array_equal(a[k], b[k]))]
print("outputs that differ between the builds:", bad or "none (bit-identical)")
PY
rm -f gpurun_out/dump_a.npz gpurun_out/dump_b.npz
