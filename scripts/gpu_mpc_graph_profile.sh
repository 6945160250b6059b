#!/bin/bash
# Per-kernel durations of ONE batched newton_solve! (graph kernel nodes profiled individually).
mkdir -p gpurun_out
CIMPC_NEWTON_HOSTLOOP=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_mpc_graph.csv python bench.py --steps 1 --warmup 3 --rollouts 16384 --mpc-rollouts 16384 --no-cpu-baseline --no-closed-loop > gpurun_out/ncu_mpc_graph.log 2>&1
tail -2 gpurun_out/ncu_mpc_graph.log | cut -c1-300
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_mpc_graph.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].split("<")[0][-40:]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[-1]) / 1e6
for k, (n, ms) in agg.items(): print(f"{k:42s} n={n:4d} total {ms:8.3f} ms")
# the last solve: kernels after the last newton_reset
last = max(i for i, r in enumerate(rows) if "newton_reset" in r[4])
seq = [(r[4].split("(")[0].split("<")[0][-24:], float(r[-1]) / 1e3) for r in rows[last:]]
print("last solve:", " ".join(f"{n.split('::')[-1][:6]}:{us:.0f}" for n, us in seq))
P
