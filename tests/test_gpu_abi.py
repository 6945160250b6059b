"""C-ABI life-cycle and misuse cases on a real device (run with `-m gpu`; `scripts/gpu_memcheck.sh` runs this file
under `compute-sanitizer --tool memcheck`).

Covers the round-1 review findings: the dense linearization must survive a simulator step on the SAME context
(`julia/CIMPCB200.jl::sim_step!` uses one context for both), a knot outside the uploaded reference must come back as
status 0 from the device entry point, and skipped subproblems must never look converged."""
import ctypes as C

import numpy as np
import pytest

from common import SIZES, load_gait, load_lin, make_batch

pytestmark = pytest.mark.gpu


def test_linearize_simstep_readback_on_one_context(cuda_device):
    """cimpc_linearize → cimpc_sim_step_batch (grows the simulator scratch) → cimpc_get_linearization →
    cimpc_linearize with the same H → solve → destroy, all on ONE context."""
    import torch
    import cimpc_b200 as cb
    capi = cb.package.capi
    robot = "quadruped"
    lin, gait = load_lin(robot), load_gait(robot)
    nq, nu, nw, nc, nb = SIZES[robot]
    opts = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=1e-8, max_ls=0, diff_sol=True)
    im = cb.ImplicitTrajectory(nq, nu, nw, nc, nb, lin["z0"], lin["th0"], kappa=float(lin["kappa"]), opts=opts)
    r0a, rza, rta = im.get_linearization()
    dev = cuda_device
    R = 96
    t = np.arange(R) % gait["u"].shape[0]
    q1 = torch.from_numpy(gait["q"][t + 1].copy()).to(dev)
    q0 = torch.from_numpy((gait["q"][t + 1] - (gait["q"][t + 1] - gait["q"][t]) / 5).copy()).to(dev)
    u = torch.from_numpy((gait["u"][t] / 5).copy()).to(dev)
    so = cb.simulator_options().to_c()
    outs = [torch.zeros((R, k), dtype=torch.float64, device=dev) for k in (nq, nc, nb)]
    st = torch.zeros(R, dtype=torch.uint8, device=dev)
    it = torch.zeros(R, dtype=torch.int32, device=dev)
    for n_step in (32, R):  # second call re-allocates the simulator scratch
        capi.check(im._ctx, im.lib.cimpc_sim_step_batch(
            im._ctx, n_step, q0.data_ptr(), q1.data_ptr(), u.data_ptr(), None, None, float(gait["mu"]),
            float(gait["h"]) / 5, C.byref(so), outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), st.data_ptr(),
            it.data_ptr(), None))
        torch.cuda.synchronize()
    assert st.cpu().numpy().mean() > 0.9
    r0b, rzb, rtb = im.get_linearization()
    assert np.array_equal(r0a, r0b) and np.array_equal(rza, rzb) and np.array_equal(rta, rtb)
    im.linearize(lin["z0"], lin["th0"], float(lin["kappa"]))  # same H: re-uses the dense arrays
    r0c, rzc, rtc = im.get_linearization()
    assert np.array_equal(rza, rzc) and np.array_equal(rta, rtc)
    knot, theta, q2 = make_batch(robot, lin, gait, 200, seed=5)
    z, dz, s2, i2 = im.solve_host(knot, theta, q2)
    assert s2.all() and np.isfinite(z).all() and np.isfinite(dz).all()
    im.close()


def test_out_of_range_knot_is_status_zero_on_device_entry(cuda_device):
    import torch
    import cimpc_b200 as cb
    robot = "quadruped"
    lin, gait = load_lin(robot), load_gait(robot)
    opts = cb.InteriorPointOptions(r_tol=1e-6, kappa_tol=1e-6, max_ls=0, diff_sol=True)
    im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], opts=opts)
    n = 700
    knot, theta, q2 = make_batch(robot, lin, gait, n, seed=11)
    H = lin["z0"].shape[0]
    bad = np.zeros(n, dtype=bool)
    bad[[0, 5, 6, 7, 300, n - 1]] = True
    skip = np.zeros(n, dtype=bool)
    skip[[1, 2, 301]] = True
    k2 = knot.copy()
    k2[bad] = H + 3
    k2[skip] = -1
    dev = cuda_device
    z, dz, st, it = im.solve_device(torch.from_numpy(k2).to(dev), torch.from_numpy(theta).to(dev), torch.from_numpy(q2).to(dev))
    torch.cuda.synchronize()
    st, it, z = st.cpu().numpy(), it.cpu().numpy(), z.cpu().numpy()
    zr, _, str_, itr = im.solve_host(knot, theta, q2)
    good = ~bad & ~skip
    assert (st[bad] == 0).all() and (it[bad] == 0).all()
    assert (st[skip] == 0).all() and (z[skip] == 0).all()
    assert np.array_equal(st[good].astype(bool), str_[good]) and np.array_equal(it[good], itr[good])
    assert np.array_equal(z[good], zr[good])
    # the host entry point rejects the same batch up front
    with pytest.raises(cb.CimpcError):
        im.solve_host(k2, theta, q2)
