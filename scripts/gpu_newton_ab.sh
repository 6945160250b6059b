#!/bin/bash
# A/B of the two specialised newton_step kernels: parity tests, per-kernel launch times of one solve (host-driven sweeps,
# under ncu) and the un-profiled solve time of each.
mkdir -p gpurun_out
if [ "$1" = "test" ]; then timeout 900 python -m pytest tests/test_gpu_newton.py tests/test_gpu_rollouts.py -m gpu -q -x 2>&1 | tail -15; fi
for k in cta warp; do
  CIMPC_NEWTON_KERNEL=$k timeout 300 python scripts/gpu_mpc_solve.py
  CIMPC_NEWTON_KERNEL=$k CIMPC_NEWTON_HOSTLOOP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/ab_launches_$k.csv python scripts/gpu_mpc_solve.py --solves 1 > gpurun_out/ab_ncu_$k.log 2>&1
  python - <<P
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/ab_launches_$k.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].split("<")[0][-40:]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[-1]) / 1e6
for kk, (n, ms) in agg.items(): print(f"{kk:42s} n={n:4d} total {ms:8.3f} ms")
last = max(i for i, r in enumerate(rows) if "newton_reset" in r[4])
seq = [(r[4].split("(")[0].split("<")[0][-24:], float(r[-1]) / 1e3) for r in rows[last:]]
print("newton_step of the last solve (us):", " ".join(f"{us:.0f}" for n, us in seq if "newton_step" in n))
P
done
