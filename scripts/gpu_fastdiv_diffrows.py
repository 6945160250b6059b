import numpy as np
a, b = np.load("gpurun_out/dump_fast.npz"), np.load("gpurun_out/dump_exact.npz")
k = "quadruped/configurationforce/"
x, y, st, it = a[k + "dz"], b[k + "dz"], a[k + "st"], a[k + "it"]
x2, y2 = x.reshape(len(x), -1), y.reshape(len(y), -1)
rows = np.flatnonzero((x2 != y2).any(1))
print("rows", rows[:30], "status", st[rows], "iters", it[rows])
for r in rows[:30]:
    d = x2[r] != y2[r]
    bothnan = np.isnan(x2[r][d]) & np.isnan(y2[r][d])
    fin = ~bothnan
    rel = np.abs(x2[r][d][fin] - y2[r][d][fin]) / np.maximum(1e-300, np.abs(y2[r][d][fin])) if fin.any() else np.array([0.0])
    print(r, "differing entries", int(d.sum()), "both NaN", int(bothnan.sum()), "max rel diff of the rest", float(np.nanmax(rel)) if rel.size else 0.0,
          "nan in fast", int(np.isnan(x2[r]).sum()), "nan in exact", int(np.isnan(y2[r]).sum()), "max|dz|", float(np.nanmax(np.abs(y2[r]))))
print("status 0 count", int((st == 0).sum()), "of", len(st))
