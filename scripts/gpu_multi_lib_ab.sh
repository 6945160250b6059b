#!/bin/bash
# Several builds of the library on the same box: per-config IP throughput of each, bitwise comparison of the solver outputs
# of each against the default build.  usage: gpu_multi_lib_ab.sh LIB1.so LIB2.so ...
mkdir -p gpurun_out
fmt='import sys,json; [print(d["robot"],d["mode"],d["subproblems"],round(d["ms"],3),"ms",round(d["subproblems_per_s"]/1e6,2),"M/s") for d in map(json.loads,sys.stdin)]'
python scripts/gpu_ip_dump.py gpurun_out/dump_a.npz > /dev/null
for rep in 1 2; do
echo "== default"; python scripts/gpu_configs.py 2>/dev/null | python -c "$fmt"
for O in "$@"; do echo "== $O"; CIMPC_B200_LIB=$O python scripts/gpu_configs.py 2>/dev/null | python -c "$fmt"; done
done
for O in "$@"; do
CIMPC_B200_LIB=$O python scripts/gpu_ip_dump.py gpurun_out/dump_b.npz > /dev/null
python - "$O" <<'PY'
import numpy as np, sys
a, b = np.load("gpurun_out/dump_a.npz"), np.load("gpurun_out/dump_b.npz")
bad = [k for k in a.files if not (np.array_equal(a[k], b[k], equal_nan=True) if a[k].dtype.kind == "f" else np.array_equal(a[k], b[k]))]
print(sys.argv[1], "outputs that differ from the default build:", bad or "none (bit-identical)")
PY
done
rm -f gpurun_out/dump_a.npz gpurun_out/dump_b.npz
