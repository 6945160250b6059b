"""Host-side bookkeeping of the closed loop (`rollout.py::MonteCarloRollouts.run`) on the CPU with stand-in kernels:
the loop that hands a whole policy interval to `cimpc_sim_steps_batch` (fused_steps = True) must record exactly what the
step-by-step loop records — `simulate!` semantics: a failed step ends the rollout (frozen configuration, skipped from then
on), `update_altitude!` takes ϕ at the FIRST step of largest impact of the interval (strict `>`, mpc_utils.jl:119), the
policy runs every N_sample steps, the last interval may be short, disturbances are indexed by simulator step.
The stand-ins follow the documented conventions of the two entry points (include/cimpc_b200.h); the GPU counterpart is
tests/test_gpu_rollouts.py::test_fused_steps_closed_loop_is_bit_identical."""
import types

import numpy as np
import pytest

torch = pytest.importorskip("torch")


class _FakeSim:
    """Toy plant: q2 = 2 q1 − q0 + h² (B u − g) + w; a step "fails" when the first coordinate leaves a band."""

    def __init__(self, nq, nu, nw, nc, nb, band):
        self.nq, self.nu, self.nw, self.nc, self.nb, self.band = nq, nu, nw, nc, nb, band
        self.calls = {"step": 0, "steps": 0}

    def _one(self, q0, q1, u, h, w, active):
        R = q0.shape[0]
        acc = torch.zeros_like(q1)
        acc[:, : self.nu] = u
        acc[:, 1] -= 9.81
        q2 = 2 * q1 - q0 + h * h * acc
        if w is not None:
            q2[:, : self.nw] += w
        gam = torch.relu(-q2[:, 1:2]).repeat(1, self.nc) * torch.arange(1, self.nc + 1, dtype=torch.float64)
        b = torch.zeros((R, self.nb), dtype=torch.float64)
        phi = q2[:, 1:2].repeat(1, self.nc) + 0.01 * torch.arange(self.nc, dtype=torch.float64)
        st = (q2[:, 0].abs() < self.band).to(torch.uint8)
        it = torch.full((R,), 7, dtype=torch.int32)
        return q2, gam, b, st, it, phi

    def step(self, q0, q1, u, mu, h, w=None, active=None, want_phi=False, opts=None, stream=None):
        self.calls["step"] += 1
        q2, gam, b, st, it, phi = self._one(q0, q1, u, h, w, active)
        if active is not None:  # skipped rollouts: q2 = q1, zero forces, status 0 (solver.py::Simulator.step)
            a = active.bool()
            q2 = torch.where(a[:, None], q2, q1)
            gam, b, phi = gam * a[:, None], b * a[:, None], phi * a[:, None]
            st, it = st * a.to(torch.uint8), it * a.to(torch.int32)
        return (q2, gam, b, st, it, phi) if want_phi else (q2, gam, b, st, it)

    def steps(self, n, q0, q1, u, mu, h, w=None, active=None, opts=None, stream=None):
        """The conventions of cimpc_sim_steps_batch: a failed step reports its forces and q2 = q_{t+1}; skipped afterwards."""
        self.calls["steps"] += 1
        R = q0.shape[0]
        run = torch.ones(R, dtype=torch.bool) if active is None else active.bool().clone()
        outs = [[] for _ in range(6)]
        qa, qb = q0, q1
        for s in range(n):
            q2, gam, b, st, it, phi = self._one(qa, qb, u, h, None if w is None else w[s], None)
            ok = st.bool() & run
            q2 = torch.where(ok[:, None], q2, qb)
            gam, b, phi = gam * run[:, None], b * run[:, None], phi * run[:, None]
            it = it * run.to(torch.int32)
            for o, v in zip(outs, (q2, gam, b, ok.to(torch.uint8), it, phi)):
                o.append(v)
            run = ok
            qa, qb = qb, q2
        return tuple(torch.stack(o) for o in outs)


class _FakeNewton:
    def __init__(self, nu):
        self.nu, self.calls = nu, []

    def solve(self, window, ref_q, ref_u, mu, h, q0, q1, warm_start=False, active=None, ref_gamma=None, ref_b=None, alt=None):
        self.calls.append((window.copy(), bool(warm_start), None if alt is None else alt.clone()))
        u = torch.tanh(q1[:, : self.nu] - q0[:, : self.nu]) + torch.from_numpy(ref_u[0])
        if alt is not None:
            u = u + alt.sum(1, keepdim=True)
        return u * active.bool()[:, None], None, None


def _make(fused, N, H_mpc=4, R=24, band=0.45, altitude=True):
    import cimpc_b200 as cb
    nq, nu, nw, nc, nb = 5, 3, 2, 2, 4
    rng = np.random.Generator(np.random.Philox(1))
    ref_q, ref_u = 0.1 * rng.standard_normal((10 + 2, nq)), 0.1 * rng.standard_normal((10, nu))
    mc = object.__new__(cb.MonteCarloRollouts)
    mc.im = types.SimpleNamespace(nq=nq, nu=nu, nw=nw, nc=nc, nb=nb)
    mc.R, mc.N, mc.H = R, N, H_mpc
    mc.h, mc.mu_mpc, mc.mu_sim = 0.05, 0.5, 1.0
    mc.ref = cb.ReferenceWindow(ref_q, ref_u, H_mpc)
    mc.newton, mc.sim = _FakeNewton(nu), _FakeSim(nq, nu, nw, nc, nb, band)
    mc.altitude_update, mc.alt_threshold = altitude, 0.02
    mc.fused_steps, mc.mpc_steps = fused, 0
    q1 = torch.from_numpy(0.2 * rng.standard_normal((R, nq)))
    v1 = torch.from_numpy(4.0 * rng.standard_normal((R, nq)))   # some rollouts drift out of the band at different steps
    return mc, q1, v1


@pytest.mark.parametrize("N,H_sim,rec", [(5, 17, 1), (3, 12, 3), (1, 6, 2)])
def test_fused_and_stepwise_loops_record_the_same(N, H_sim, rec):
    outs = []
    for fused in (True, False):
        mc, q1, v1 = _make(fused, N)
        dist = lambda t: np.array([1e-3 * t, -2e-3 * (t % 3)])  # noqa: E731  w of simulator step t
        out = mc.run(q1, v1, H_sim, record_every=rec, dist=dist)
        n_policy = -(-H_sim // N)
        assert mc.mpc_steps == n_policy == len(mc.newton.calls)
        assert [c[1] for c in mc.newton.calls] == [False] + [True] * (n_policy - 1)          # warm start after the first call
        assert (mc.sim.calls["steps"], mc.sim.calls["step"]) == ((n_policy, 0) if fused else (0, H_sim))
        outs.append((out, mc.newton.calls))
    (a, ca), (b, cb_) = outs
    for k in ("q", "u", "gamma", "b", "status", "failed_at", "alt"):
        assert torch.equal(a[k], b[k]), k
    for x, y in zip(ca, cb_):                                                                # the policy saw the same inputs
        assert np.array_equal(x[0], y[0]) and torch.equal(x[2], y[2])
    out = a
    failed = out["failed_at"].numpy()
    assert (failed > 0).any() and (failed == 0).any(), "the toy band must end some rollouts and keep others"
    assert np.array_equal(out["status"].numpy(), failed == 0)
    # a rollout that failed at step t is frozen from q_{t+1} on (records every `rec` steps: index k ↔ step (k − 1)·rec)
    q = out["q"].numpy()
    for r in np.flatnonzero(failed > 0):
        k0 = -(-(int(failed[r]) + 1) // rec) + 1 if rec > 1 else int(failed[r]) + 1
        tail = q[min(k0, q.shape[0] - 1):, r]
        assert np.all(tail == tail[0])


def test_altitude_is_taken_at_the_first_step_of_largest_impact():
    mc, q1, v1 = _make(True, 4, altitude=True, band=10.0)
    out = mc.run(q1, v1, 12)
    alts = [c[2] for c in mc.newton.calls]
    assert torch.all(alts[0] == 0)                       # policy.jl:111-115: no update before the first interval has run
    assert any(bool((a != 0).any()) for a in alts[1:])   # impacts above the threshold moved some altitudes
    assert torch.equal(out["alt"], alts[-1]) or out["alt"].shape == alts[-1].shape
