#!/bin/bash
# One ncu --set full capture per kernel of the path (config 5 of BASELINE.json asks for a roofline report per kernel).
mkdir -p gpurun_out
cat > /tmp/one_cfg.py <<'PY'
import os, sys, numpy as np, torch
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import SIZES, load_gait, load_lin, make_batch
robot, mode, H, n = "centroidal_quadruped", "configuration", 20, 262144
lin, gait = load_lin(robot), load_gait(robot)
im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], kappa=float(lin["kappa"]), mode=mode,
                           opts=cb.InteriorPointOptions(diff_sol=True, r_tol=1e-8, kappa_tol=2e-4))   # linearize_kernel + prep_kernel
knot, theta, q2 = make_batch(robot, lin, gait, n, seed=1)
R = n // H
stage = (np.arange(n) // R).astype(np.int32) % lin["z0"].shape[0]
nq = SIZES[robot][0]
theta = theta - lin["th0"][knot] + lin["th0"][stage]; q2 = q2 - lin["z0"][knot, :nq] + lin["z0"][stage, :nq]
dev = torch.device("cuda:0")
kd, td, qd = torch.from_numpy(stage).to(dev), torch.from_numpy(np.ascontiguousarray(theta)).to(dev), torch.from_numpy(np.ascontiguousarray(q2)).to(dev)
out = im.solve_device(kd, td, qd); im.solve_device(kd, td, qd, out=out); torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none -k regex:'ip_solve|prep_kernel|linearize' -c 4 -f -o gpurun_out/prof_centroidal python /tmp/one_cfg.py > gpurun_out/ncu_centroidal.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'newton_step|newton_reset|newton_finish' -c 4 -f -o gpurun_out/prof_newton2 python bench.py --steps 1 --warmup 1 --rollouts 4096 --mpc-rollouts 16384 --no-cpu-baseline --no-closed-loop > gpurun_out/ncu_newton2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:sim_step -c 2 -f -o gpurun_out/prof_sim2 python scripts/gpu_sim_profile.py > gpurun_out/ncu_sim2.log 2>&1
ls -la gpurun_out/*.ncu-rep
