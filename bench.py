#!/usr/bin/env python
"""bench.py — contact-subproblems/s of the batched linearized interior-point sweep.

    python bench.py --gpus 1 --steps 5 --warmup 3              # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps 2 ...    # the reference algorithm on host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one `implicit_dynamics!` sweep (src/controller/implicit_dynamics.jl:156-192) over the
whole Monte-Carlo batch: quadruped (`examples/quadruped/monte_carlo.jl:20-58`: gait2, H_mpc = 10,
κ = 1e-4, IP r_tol = κ_tol = 1e-4, max_iter = 100, diff_sol), 65 536 rollouts × 10 stages = 655 360
cold-started subproblems PER GPU, one kernel launch.  Weak scaling: every rank owns its own 64k
rollouts (independent, no data-path collective); rank 0 only gathers iteration statistics.

`value`   subproblems/s with inputs resident in HBM (device entry point of the C ABI).
`e2e`     the same sweep through `cimpc_ip_solve_batch_host` with HOST (pinned) buffers: H2D of
          (knot, θ, q2) and D2H of (z, δz, status, iters) inside the timed region.
`roofline` HBM roofline of ip_solve_kernel from its algorithmic bytes (3 125 B per
          quadruped subproblem) — structurally a few % because the path is fp64-issue bound; the
          `fp64` object gives the companion fp64-FMA fraction.
`cpu_baseline` the C restatement of the reference algorithm (oracle/c, MGS-QR, all host threads).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ROBOT = "quadruped"
MODE = "configuration"
H_MPC = 10
ROLLOUTS_PER_GPU = 65536
IP_KW = dict(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, diff_sol=True)  # monte_carlo.jl:51-58
METRIC = "contact_subproblems_per_sec"
UNIT = "subproblems/s"


def algorithmic_bytes(nq, nu, nw, nc, nb, mode):
    """SURVEY.md §8d / BASELINE.md §3: read θ, q2 guess, alt; write d, δq0, δq1, δu1, status, iters."""
    nth = 2 * nq + nu + nw + 2
    nd = nq if mode == "configuration" else nq + nc + nb
    return 8 * (nth + nq + nc) + 8 * (nd + nd * (2 * nq + nu)) + 5


def algorithmic_flops(nq, nu, nc, nb, mean_iters):
    """BASELINE.md §3 (reference algorithm's count: MGS 2ny³ per factorisation)."""
    nx, ny = nq, 2 * nc + nb
    f_it = 2 * ny ** 3 + ny ** 2 + 4 * (2 * nx * ny + 1.5 * ny ** 2 + nx ** 2) + 6 * (nx ** 2 + 2 * nx * ny + ny ** 2)
    f_d = 2 * ny ** 3 + (2 * nq + nu) * 2 * (2 * nx * ny + 1.5 * ny ** 2 + nx ** 2)
    return mean_iters * f_it + f_d


def build_workload(rank: int, rollouts: int):
    """Stage-major batch: subproblem (stage i, rollout r) uses knot i (all rollouts share the window,
    policy.jl:100-107, 131) with rollout-specific perturbed θ and cold-start q2."""
    from common import SIZES, load_gait, load_lin, make_batch
    lin, gait = load_lin(ROBOT), load_gait(ROBOT)
    n = rollouts * H_MPC
    knot, theta, q2 = make_batch(ROBOT, lin, gait, n, seed=100 + rank)
    # make_batch cycles knots 0..H_ref-1; restrict to the MPC window 0..H_MPC-1, stage-major
    nq = SIZES[ROBOT][0]
    stage = (np.arange(n) // rollouts).astype(np.int32)
    theta = theta - lin["th0"][knot] + lin["th0"][stage]
    q2 = q2 - lin["z0"][knot, :nq] + lin["z0"][stage, :nq]
    return lin, stage, np.ascontiguousarray(theta), np.ascontiguousarray(q2)


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md): an NVML polling thread
    (≈ 2 ms period — the timed region of a default run is only tens of ms); `nvidia-smi -lms` as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []  # (time, sm_mhz, max_mhz, set(reasons))
        self.index = index
        self.proc = None
        self.nvml = None
        self._stop = False
        self.source = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES remaps torch's index: resolve through the PCI bus id
            import torch
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(self.index), "pci_bus_id") else None
            h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if int(pynvml.nvmlDeviceGetPciInfo(hi).bus) == int(bus):
                        h = hi
                        break
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml = (pynvml, h)
            self.source = "nvml"
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nvml
        R = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
             "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
             "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
             "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        while not self._stop:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((time.time(), sm, mx, {k for k, v in R.items() if bits & v}))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.rows.append((time.time(), float(f[0]), float(f[1]),
                                  {nm for nm, v in zip(names, f[3:7]) if v.lower().startswith("active")}))
            except Exception:
                continue

    def stop(self, t0, t1):
        self._stop = True
        if self.proc is not None:
            self.proc.terminate()
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml / nvidia-smi unavailable"], "samples": 0}
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        note = None
        if not inside:  # region shorter than the sampling period: take the nearest samples
            inside = sorted(self.rows, key=lambda r: min(abs(r[0] - t0), abs(r[0] - t1)))[:3]
            note = "no sample inside the timed region; nearest samples used"
        sm = [r[1] for r in inside]
        reasons = set().union(*[r[3] for r in inside]) if inside else set()
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": inside[0][2] if inside else None,
               "reasons": sorted(reasons), "samples": len(sm), "source": self.source}
        if note:
            out["note"] = note
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def fp64_peak():
    """fp64 FMA peak measured by profiles/tools/fp64_peak.cu (committed summary), else nominal."""
    p = os.path.join(ROOT, "profiles", "fp64_peak.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["dfma_tflops"]), "measured (profiles/fp64_peak.json)"
    return 37.2, "nominal 148 SM x 64 DFMA/clk x 1.965 GHz"


def cpu_baseline_run(lin, knot, theta, q2, sample: int, threads: int = 0):
    from common import SIZES
    from oracle.c_oracle import COracle
    from oracle.ip import IPOptions
    co = COracle(*SIZES[ROBOT], lin, mode=MODE, solver="mgs")
    o = IPOptions(**IP_KW)
    co.solve(knot[:2048], theta[:2048], q2[:2048], o, threads=threads)  # warm-up (page-in, thread pool)
    t0 = time.perf_counter()
    z, dz, st, it = co.solve(knot[:sample], theta[:sample], q2[:sample], o, threads=threads)
    dt = time.perf_counter() - t0
    return sample / dt, co.max_threads if threads <= 0 else threads, dt, float(it.mean()), float(st.mean())


def run_reference(args):
    """`--impl reference`: the reference's algorithm (C restatement: explicit Dx⁻¹, MGS-QR per
    iteration, all nθ sensitivity columns) on the host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lin, knot, theta, q2 = build_workload(0, ROLLOUTS_PER_GPU // 8)
    n = knot.shape[0]  # 81 920 subproblems per step (1/8 of the per-GPU batch)
    from common import SIZES
    from oracle.c_oracle import COracle
    from oracle.ip import IPOptions
    co = COracle(*SIZES[ROBOT], lin, mode=MODE, solver="mgs")
    o = IPOptions(**IP_KW)
    for _ in range(max(args.warmup, 1)):
        co.solve(knot, theta, q2, o)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        z, dz, st, it = co.solve(knot, theta, q2, o)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    cores = co.max_threads
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"quadruped flat gait2 H_mpc=10 linearized IP sweep, bounded sample of "
                               f"{n} subproblems/step of the 655360-subproblem batch", "mode": MODE, **IP_KW},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} subproblems x {args.steps} steps; C restatement of the Julia CPU path "
                                   f"(oracle/c/ip_oracle.c, OpenMP over problems)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mean_ip_iterations": float(it.mean()), "converged_frac": float(st.mean()),
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rollouts", type=int, default=ROLLOUTS_PER_GPU, help="rollouts per GPU (default 65536)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mpc", action="store_true", help="skip the batched newton_solve! (MPC steps/s) leg")
    ap.add_argument("--mpc-rollouts", type=int, default=16384)
    ap.add_argument("--no-closed-loop", action="store_true", help="skip the on-device Monte-Carlo closed-loop leg")
    ap.add_argument("--closed-loop-groups", type=int, default=2, help="independent parts (streams) of the closed-loop leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import cimpc_b200 as cb
    from common import SIZES

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    nq, nu, nw, nc, nb = SIZES[ROBOT]
    lin, knot, theta, q2 = build_workload(rank, args.rollouts)
    n = knot.shape[0]
    opts = cb.InteriorPointOptions(**IP_KW)
    im = cb.ImplicitTrajectory(nq, nu, nw, nc, nb, lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"],
                               mode=MODE, opts=opts, device=local)

    # ---- device-resident leg -------------------------------------------------------------
    kd = torch.from_numpy(knot).to(dev)
    td = torch.from_numpy(theta).to(dev)
    qd = torch.from_numpy(q2).to(dev)
    outb = im.solve_device(kd, td, qd)
    for _ in range(args.warmup):
        im.solve_device(kd, td, qd, out=outb)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = im.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    torch.cuda.synchronize()
    t_wall0 = time.time()
    ev[0].record()
    for s in range(args.steps):
        im.solve_device(kd, td, qd, out=outb)
        ev[s + 1].record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    if dist:
        dist.barrier()
    launches = im.launch_count - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    st = outb[2]
    it = outb[3]
    stats = torch.stack([st.double().mean(), it.double().mean(), it.double().max()])
    if dist:
        gathered = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats)
        stats = torch.stack(gathered).mean(0)
    conv_frac, mean_it, max_it = [float(x) for x in stats.cpu()]

    # ---- end-to-end leg: host (pinned) buffers through cimpc_ip_solve_batch_host ---------------
    e2e_steps = max(2, min(args.steps, 3))
    knot_h = torch.from_numpy(knot).pin_memory()
    theta_h = torch.from_numpy(theta).pin_memory()
    q2_h = torch.from_numpy(q2).pin_memory()
    z_h = torch.empty((n, im.nz), dtype=torch.float64).pin_memory()
    dz_h = torch.empty((n, im.ncol, im.nd), dtype=torch.float64).pin_memory()
    st_h = torch.empty(n, dtype=torch.uint8).pin_memory()
    it_h = torch.empty(n, dtype=torch.int32).pin_memory()
    h2d = knot_h.numel() * 4 + theta_h.numel() * 8 + q2_h.numel() * 8
    d2h = z_h.numel() * 8 + dz_h.numel() * 8 + st_h.numel() + it_h.numel() * 4

    def e2e_once():
        im.solve_host_into(knot_h.numpy(), theta_h.numpy(), q2_h.numpy(), None, z_h.numpy(), dz_h.numpy(),
                           st_h.numpy(), it_h.numpy())

    e2e_once()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_once()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if dist:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e_ok = bool(np.array_equal(z_h.numpy(), outb[0].cpu().numpy()))

    # ---- MPC-step leg: batched newton_solve! (cold start, one MPC step of every rollout) --------------
    mpc = None
    if not args.no_mpc:
        from common import load_gait
        gait = load_gait(ROBOT)
        Rm = min(args.rollouts, args.mpc_rollouts)
        oq = np.tile(1e-2 * np.array([1.0, 0.02, 0.25] + [0.75] * (nq - 3)), (H_MPC, 1))  # monte_carlo.jl:33-37
        ou = np.tile(3e-2 * np.ones(nu), (H_MPC, 1))
        newton = cb.Newton(im, H_MPC, Rm, oq, ou, 1.0e-4, cb.NewtonOptions(r_tol=3e-4, max_iter=5), ip_opts=opts)
        rng = np.random.Generator(np.random.Philox(1000 + rank))
        q0m = torch.from_numpy(np.tile(gait["q"][0], (Rm, 1))).to(dev)
        q1m = torch.from_numpy(gait["q"][1] + 0.01 * rng.standard_normal((Rm, nq))).to(dev)
        win = np.arange(H_MPC + 2, dtype=np.int32)
        newton.solve(win, gait["q"][:H_MPC + 2], gait["u"][:H_MPC], gait["mu"], gait["h"], q0m, q1m)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        m_steps = 3
        l0 = im.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(m_steps):
            um, _, infom = newton.solve(win, gait["q"][:H_MPC + 2], gait["u"][:H_MPC], gait["mu"], gait["h"], q0m, q1m)
        e1.record()
        torch.cuda.synchronize()
        tm = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if dist:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        inf = infom.double().mean(0).cpu().numpy()
        mpc = {"value": Rm * world * m_steps / (float(tm.item()) * 1e-3), "unit": "MPC steps/s",
               "rollouts_per_gpu": Rm, "steps": m_steps, "ms_per_mpc_step_batch": float(tm.item()) / m_steps,
               "mean_newton_iterations": float(inf[0]), "mean_implicit_dynamics_sweeps": float(inf[1]),
               "converged_frac": float(inf[2]), "sweeps_per_call": newton.last_sweeps,
               # one CUDA-graph launch per MPC step; its kernel nodes: reset + finish + 3 per sweep slot (unrolled to the
               # reference's sweep bound 1 + 8·max_iter; slots after the last active rollout exit immediately)
               "host_launch_calls_per_mpc_step": 1, "host_syncs_per_mpc_step": 0,
               "gpu_kernels_doing_work_per_mpc_step": 2 + 3 * int(newton.last_sweeps),
               "gpu_launches": int(im.launch_count - l0),
               "config": "newton_solve! cold start, quadruped H_mpc=10, r_tol=3e-4, max_iter=5 (monte_carlo.jl:44-48), "
                         "q1 = reference + N(0, 0.01^2)"}

    # ---- closed-loop leg: the Monte-Carlo workload itself (policy + simulator on the device) ----------
    closed = None
    if not args.no_mpc and not args.no_closed_loop:
        Rc = min(args.rollouts, args.mpc_rollouts)
        N_sample = 5
        # the rollouts run as `--closed-loop-groups` independent parts (own stream + host thread each): one part's
        # simulator latency tail overlaps the other parts' work; per-rollout results do not depend on the split
        def make_im():
            return cb.ImplicitTrajectory(*SIZES[ROBOT], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode=MODE,
                                         opts=opts, device=dev.index or 0)
        mc = cb.GroupedRollouts(make_im, Rc, args.closed_loop_groups, gait["q"], gait["u"], gait["mu"], 1.0, gait["h"],
                                H_mpc=H_MPC, N_sample=N_sample, obj_q=oq, obj_u=ou, kappa=1.0e-4,
                                newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5))
        lo, hi = cb.shard_rollouts(Rc * world, world, rank)
        qinit = cb.quadruped_initial_configurations(Rc * world, seed=100)[lo:hi]
        q1c = torch.from_numpy(qinit).to(dev)
        v1c = torch.from_numpy(np.tile((gait["q"][1] - gait["q"][0]) / gait["h"], (Rc, 1))).to(dev)
        H_sim_c = 10 * N_sample  # 10 MPC steps of every rollout
        mc.run(q1c, v1c, N_sample)  # warm-up: one MPC step + N_sample simulator steps
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        l0 = im.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res_c = mc.run(q1c, v1c, H_sim_c, record_every=N_sample)
        e1.record()
        torch.cuda.synchronize()
        tc = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        okc = res_c["status"].double().mean().reshape(1)
        if dist:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            dist.all_reduce(okc, op=dist.ReduceOp.SUM)
            okc /= world
            # final gather of the (down-sampled: every N_sample-th step) configurations to rank 0 over NCCL
            qg = cb.gather_rollout_results(res_c["q"].permute(1, 0, 2).contiguous(), Rc * world)
            gathered = None if qg is None else list(qg.shape)
        else:
            gathered = list(res_c["q"].permute(1, 0, 2).shape)
        closed = {"value": Rc * world * mc.mpc_steps / (float(tc.item()) * 1e-3), "unit": "MPC steps/s (closed loop)",
                  "rollouts_per_gpu": Rc, "sim_steps": H_sim_c, "mpc_steps_per_rollout": mc.mpc_steps,
                  "ms_total": float(tc.item()), "sim_ok_frac": float(okc.item()), "groups": args.closed_loop_groups,
                  "gathered_trajectory_shape": gathered,
                  "config": "examples/quadruped/monte_carlo.jl: initial configurations from the conf_min/conf_max box "
                            "(Philox seed 100), policy every 5 simulator steps, nonlinear simulator step on the device; "
                            "trajectories recorded every 5th step and gathered to rank 0"}

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    value = n * world * args.steps / (total_ms_max * 1e-3)
    kernel_ms = float(np.mean(per_launch_ms))
    peaks, peak_kind = measured_peaks()
    B = algorithmic_bytes(nq, nu, nw, nc, nb, MODE)
    achieved = n * B / (kernel_ms * 1e-3) / 1e9
    fl = algorithmic_flops(nq, nu, nc, nb, mean_it)
    f_peak, f_kind = fp64_peak()
    f_ach = n * fl / (kernel_ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ip_kernel_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            # measured on an 81 920-subproblem launch; DRAM traffic is per-subproblem streaming, so scale to n
            traffic = json.load(f).get("dram_bytes_per_subproblem") * n

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"quadruped flat gait2, H_mpc={H_MPC}, {args.rollouts} Monte-Carlo rollouts per GPU = "
                               f"{n} cold-started linearized IP subproblems per step per GPU (one implicit_dynamics! sweep)",
                   "mode": MODE, **IP_KW, "l2": "inputs+outputs per step (2.2 GB) exceed the 126 MB L2",
                   "parallelism": f"rollout-sharded x{world}, no data-path collective"},
        "mean_ip_iterations": mean_it, "max_ip_iterations": max_it, "converged_frac": conv_frac,
        "clocks": clocks,
        "gpu_launches": int(launches),
        "e2e": {"value": n * world * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "matches_device_leg": e2e_ok,
                "api": "cimpc_ip_solve_batch_host (pinned host buffers, chunk-pipelined H2D / kernel / D2H)"},
        "roofline": {"bound": "hbm", "kernel": "ip_solve_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                     "peak_source": f"MEASURED_PEAKS.json ({peak_kind})", "algorithmic_bytes_per_subproblem": B,
                     "kernel_ms": kernel_ms,
                     "note": "fp64-issue bound path: HBM fraction is structurally small (SURVEY.md §8d); see fp64"},
        "fp64": {"achieved": f_ach, "peak": f_peak, "unit": "TFLOP/s", "frac": f_ach / f_peak,
                 "flops_per_subproblem": fl, "peak_source": f_kind,
                 "note": "flops counted as the REFERENCE algorithm would spend them (BASELINE.md §3)"},
    }
    if mpc is not None:
        out["mpc_steps"] = mpc
    if closed is not None:
        out["closed_loop"] = closed
    if not args.no_cpu_baseline and world == 1:
        sample = min(n, 262144)
        v, cores, dt, mit, conv = cpu_baseline_run(lin, knot, theta, q2, sample)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"first {sample} subproblems of the same batch, {dt:.1f} s wall; C restatement "
                                         f"of the Julia CPU path (explicit Dx^-1, MGS-QR per iteration, all n_theta "
                                         f"sensitivity columns), OpenMP over problems",
                               "mean_ip_iterations": mit, "converged_frac": conv}
    print(json.dumps(out))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
