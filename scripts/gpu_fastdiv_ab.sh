#!/bin/bash
# The branch-free reciprocal / division of csrc/fastdiv.cuh (default build) against a -DCIMPC_IP_FASTDIV=0 build
# (lib/libcimpc_b200_exactdiv.so: CIMPC_BUILD_TAG=exactdiv CIMPC_EXTRA_NVCC_FLAGS=-DCIMPC_IP_FASTDIV=0 python build.py):
# operation-level accuracy, bitwise comparison of the solver outputs on every robot, device-resident throughput, MPC step.
mkdir -p gpurun_out
E=contactimplicitmpc.jl_b200/lib/libcimpc_b200_exactdiv.so
./profiles/tools/rcp_check | tee gpurun_out/rcp_check.json
python scripts/gpu_ip_dump.py gpurun_out/dump_fast.npz
CIMPC_B200_LIB=$E python scripts/gpu_ip_dump.py gpurun_out/dump_exact.npz
python - <<'PY'
import numpy as np, json
a, b = np.load("gpurun_out/dump_fast.npz"), np.load("gpurun_out/dump_exact.npz")
res = {}
for k in a.files:
    x, y = a[k], b[k]
    eq = np.array_equal(x, y, equal_nan=True) if x.dtype.kind == "f" else np.array_equal(x, y)
    res[k] = {"bitwise_equal": bool(eq), "n_differing_rows": int((x.reshape(len(x), -1) != y.reshape(len(y), -1)).any(1).sum()), "rows": int(len(x))}
print(json.dumps(res, indent=1))
json.dump(res, open("gpurun_out/fastdiv_bitwise.json", "w"), indent=1)
PY
rm -f gpurun_out/dump_fast.npz gpurun_out/dump_exact.npz
fmt='import sys,json; [print(d["robot"],d["mode"],d["subproblems"],round(d["ms"],3),"ms",round(d["subproblems_per_s"]/1e6,2),"M/s") for d in map(json.loads,sys.stdin)]'
echo "== exact"; CIMPC_B200_LIB=$E python scripts/gpu_configs.py 2>/dev/null | python -c "$fmt"
echo "== fast (default)"; python scripts/gpu_configs.py 2>/dev/null | python -c "$fmt"
echo "== mpc exact"; CIMPC_B200_LIB=$E python scripts/gpu_mpc_solve.py --solves 4 | tail -2
echo "== mpc fast (default)"; python scripts/gpu_mpc_solve.py --solves 4 | tail -2
