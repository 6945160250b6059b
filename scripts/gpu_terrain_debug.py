import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cimpc_b200 as cb
from common import SIZES, load_gait
from oracle.ip import IPOptions
from oracle.residual import get_residual
from oracle.simulator import nonlinear_ip_solve
robot = sys.argv[1] if len(sys.argv) > 1 else "flamingo"; tag = robot + "_piecewise"
res = get_residual(tag); m = res.model
gait = load_gait(robot); H = gait["u"].shape[0]; h_sim = gait["h"] / 5
rng = np.random.default_rng(51); R = 40
t = rng.integers(0, H, R)
q1 = gait["q"][t + 1] + 0.002 * rng.standard_normal((R, m.nq))
v = (gait["q"][t + 1] - gait["q"][t]) / gait["h"]
x_new = rng.uniform(-0.2, 2.8, R); x_new[:6] = [0.45, 0.5, 0.55, 1.95, 2.0, 2.05]
q1[:, 0] = x_new; q1[:, 1] += np.array([res.terrain.height(x) for x in x_new])
q0 = q1 - h_sim * v * (1 + 0.05 * rng.standard_normal((R, 1)))
u = gait["u"][t] / 5 * (1 + 0.1 * rng.standard_normal((R, m.nu)))
mu = m.mu_world
o = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=1e-8, max_ls=0, eps_min=0.25, undercut=float("inf"), gamma_reg=0.1)
sim = cb.Simulator(*SIZES[robot], opts=o, model=tag)
dev = torch.device("cuda", 0)
out = sim.step(torch.from_numpy(q0).to(dev), torch.from_numpy(q1).to(dev), torch.from_numpy(u).to(dev), mu, h_sim, want_phi=True)
torch.cuda.synchronize()
q2, gam, b, st, it, phi = [x.cpu().numpy() for x in out]
oo = IPOptions(r_tol=1e-8, kappa_tol=1e-8, max_ls=0, eps_min=0.25, undercut=np.inf, gamma_reg=0.1)
i = res.idx
for r in range(R):
    z = np.ones(i.nz); z[i.q2] = q1[r]
    th = np.concatenate([q0[r], q1[r], u[r], np.zeros(m.nw), [mu], [h_sim]])
    ok, zo, ito = nonlinear_ip_solve(res, z, th, oo)
    dq, dg, db = np.abs(q2[r] - zo[i.q2]).max(), np.abs(gam[r] - zo[i.g1]).max(), np.abs(b[r] - zo[i.b1]).max()
    zd = zo.copy(); zd[i.q2] = q2[r]; zd[i.g1] = gam[r]; zd[i.b1] = b[r]; zd[i.s1] = phi[r]
    rr = res.r(zd, th, 0.0)
    px = res._px(q2[r])
    if max(dq, dg, db) > 1e-6:
        print(f"r={r} x={x_new[r]:.3f} st={st[r]} it={it[r]}/{ito} dq={dq:.2e} dg={dg:.2e} db={db:.2e} |gam|={np.abs(zo[i.g1]).max():.2f} "
              f"r_dyn={np.abs(rr[i.dyn]).max():.1e} r_imp={np.abs(rr[i.imp]).max():.1e} px={np.round(px,3)} slopes={[round(res.terrain.slope(x),3) for x in px]}")
print("done")
