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

# BASELINE.json `configs` 2-5 as IP-sweep workloads (config 1 is the CPU-only hopper call, tests/test_oracle.py).
# rollouts × H_mpc subproblems per step and GPU; `alt`: non-zero altitude offsets on the impact rows (config 3:
# examples/flamingo/piecewise.jl runs the flat-linearized policy with measured altitudes).
CONFIGS = {
    2: dict(robot="quadruped", mode="configuration", H=10, rollouts=4096, alt=False,
            ip=dict(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, diff_sol=True),
            name="quadruped flat gait2, H_mpc=10, 4096 Monte-Carlo rollouts (examples/quadruped/monte_carlo.jl:27-58)"),
    3: dict(robot="flamingo", mode="configurationforce", H=15, rollouts=1093, alt=True,
            ip=dict(r_tol=1e-8, kappa_tol=1e-4, max_iter=100, diff_sol=True, undercut=5.0, gamma_reg=0.1),
            name="flamingo gait_forward_36_4, H_mpc=15, :configurationforce, 16 395-problem IP batch with non-zero altitude "
                 "offsets (examples/flamingo/piecewise.jl:24-48; ip_opts defaults of ci_mpc_policy, policy.jl:54-61)"),
    4: dict(robot="quadruped", mode="configuration", H=10, rollouts=65536, alt=False,
            ip=dict(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, diff_sol=True),
            name="quadruped flat gait2, H_mpc=10, 65536 Monte-Carlo rollouts per GPU"),
    5: dict(robot="centroidal_quadruped", mode="configuration", H=20, rollouts=13108, alt=False,
            ip=dict(r_tol=1e-4, kappa_tol=2e-4, max_iter=100, diff_sol=True, undercut=5.0),
            name="centroidal_quadruped inplace_trot_v4, H_mpc=20, 13 108 rollouts = 262 160 subproblems ('256k batch') per GPU "
                 "(examples/centroidal_quadruped/flat_trot.jl:31-62); the payload only changes the simulated plant, "
                 "not this kernel (cimpc_create_named)"),
}


def executed_flops(nq, nu, nw, nc, nb, mode, mean_iters, evals_per_iter=1.0):
    """Flops THIS implementation executes per subproblem (csrc/ip_kernel.cuh; FMA = 2 flops, useful lanes only): per
    iteration the assembly of the reduced (nc+nb)² block, its in-place Gauss-Jordan inversion (ψ passenger rows
    included), CAi·r and Ai·r, two inverse applications with the ψ recovery, Δx, and `evals_per_iter` candidate
    residuals; once per subproblem the θ prologue, the first residual, and the sensitivities (a second inversion,
    P1 = AiB S⁻¹, and the products with the 2nq+nu consumed columns of W)."""
    nx, ny, nr, npsi = nq, 2 * nc + nb, nc + nb, nc
    nth, ncol, nfr = 2 * nq + nu + nw + 2, 2 * nq + nu, 1 + nb // nc
    nyd = nr if mode == "configurationforce" else 0
    res = 2 * (nx + ny) ** 2
    per_it = 3 * nr ** 2 + 2 * nr ** 3 + 2 * nr ** 2 * npsi + 4 * nx * ny + 2 * (2 * nr ** 2 + 2 * nfr * npsi) + 2 * nx * nr \
        + evals_per_iter * res
    once = 2 * nth * (nx + ny) + res + 9 * nr ** 2 + 2 * nr ** 3 + 2 * nx * nr ** 2 + 2 * ncol * nr * (nx + nyd)
    return mean_iters * per_it + once


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


def build_workload(rank: int, rollouts: int, cfg=None):
    """Stage-major batch: subproblem (stage i, rollout r) uses knot i (all rollouts share the window,
    policy.jl:100-107, 131) with rollout-specific perturbed θ and cold-start q2.
    Returns lin, knot, θ, q2, alt (None unless the config asks for altitude offsets)."""
    from common import SIZES, load_gait, load_lin, make_batch
    cfg = cfg or CONFIGS[4]
    robot, H = cfg["robot"], cfg["H"]
    lin, gait = load_lin(robot), load_gait(robot)
    n = rollouts * H
    knot, theta, q2 = make_batch(robot, lin, gait, n, seed=100 + rank)
    # make_batch cycles knots 0..H_ref-1; restrict to the MPC window 0..H-1, stage-major
    nq, nc = SIZES[robot][0], SIZES[robot][3]
    stage = (np.arange(n) // rollouts).astype(np.int32)
    theta = theta - lin["th0"][knot] + lin["th0"][stage]
    q2 = q2 - lin["z0"][knot, :nq] + lin["z0"][stage, :nq]
    alt = None
    if cfg.get("alt"):  # piecewise terrain: steps of up to 2 cm under the feet, one altitude vector per rollout
        rng = np.random.Generator(np.random.Philox(7000 + rank))
        alt = np.ascontiguousarray(np.tile(0.02 * rng.random((rollouts, nc)), (H, 1)))
    return lin, stage, np.ascontiguousarray(theta), np.ascontiguousarray(q2), alt


def ip_sweep_leg(cb, torch, dev, cfg, rank, steps, warmup, rollouts=None):
    """Device-resident throughput of one `implicit_dynamics!` sweep of a BASELINE config: value, ms, both rooflines."""
    from common import SIZES
    robot, mode = cfg["robot"], cfg["mode"]
    nq, nu, nw, nc, nb = SIZES[robot]
    lin, knot, theta, q2, alt = build_workload(rank, rollouts or cfg["rollouts"], cfg)
    n = knot.shape[0]
    im = cb.ImplicitTrajectory(nq, nu, nw, nc, nb, lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode=mode,
                               opts=cb.InteriorPointOptions(**cfg["ip"]), device=dev.index or 0)
    kd, td, qd = torch.from_numpy(knot).to(dev), torch.from_numpy(theta).to(dev), torch.from_numpy(q2).to(dev)
    ad = torch.from_numpy(alt).to(dev) if alt is not None else None
    outb = im.solve_device(kd, td, qd, alt=ad)
    for _ in range(warmup):
        im.solve_device(kd, td, qd, alt=ad, out=outb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        im.solve_device(kd, td, qd, alt=ad, out=outb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    mean_it = float(outb[3].double().mean().item())
    peaks, peak_kind = measured_peaks()
    f_peak, f_kind = fp64_peak()
    B = algorithmic_bytes(nq, nu, nw, nc, nb, mode)
    fl_ref = algorithmic_flops(nq, nu, nc, nb, mean_it)
    fl_exe = executed_flops(nq, nu, nw, nc, nb, mode, mean_it)
    gbs = n * B / (ms * 1e-3) / 1e9
    res = {"workload": cfg["name"], "robot": robot, "mode": mode, "H_mpc": cfg["H"], "subproblems_per_step": int(n),
           "value": n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "mean_ip_iterations": mean_it,
           "converged_frac": float(outb[2].double().mean().item()), "lanes_per_subproblem": im.group,
           "l2": "inputs + outputs per step exceed the 126 MB L2" if n * B > 126e6 else
                 f"inputs + outputs per step = {n * B / 1e6:.0f} MB < 126 MB L2: resident after warm-up (a {cfg['rollouts']}-rollout "
                 "batch is this small by definition of the config)",
           "roofline": {"bound": "hbm", "kernel": "ip_solve_kernel", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": gbs / peaks["hbm_gbs"], "algorithmic_bytes_per_subproblem": B, "traffic": None},
           "fp64": {"peak": f_peak, "unit": "TFLOP/s",
                    "executed": {"achieved": n * fl_exe / (ms * 1e-3) / 1e12, "frac": n * fl_exe / (ms * 1e-3) / 1e12 / f_peak,
                                 "flops_per_subproblem": fl_exe},
                    "reference_count": {"achieved": n * fl_ref / (ms * 1e-3) / 1e12,
                                        "frac": n * fl_ref / (ms * 1e-3) / 1e12 / f_peak, "flops_per_subproblem": fl_ref}}}
    im.close()
    return res


def pin_to_gpu_numa(local: int):
    """Bind this rank's host threads (and therefore its first-touch pinned allocations) to the NUMA node of its GPU
    (/sys/bus/pci/devices/<bus id>/local_cpulist); best effort, reported in the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        dev_id = getattr(torch.cuda.get_device_properties(local), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0"
        with open(os.path.join(path, "local_cpulist")) as f:
            txt = f.read().strip()
        with open(os.path.join(path, "numa_node")) as f:
            node = int(f.read().strip())
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception as e:  # noqa: BLE001
        return {"numa_node": None, "error": type(e).__name__}


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


def cpu_baseline_run(lin, knot, theta, q2, sample: int, threads: int = 0, cfg=None, alt=None):
    from common import SIZES
    from oracle.c_oracle import COracle
    from oracle.ip import IPOptions
    cfg = cfg or CONFIGS[4]
    co = COracle(*SIZES[cfg["robot"]], lin, mode=cfg["mode"], solver="mgs")
    o = IPOptions(**cfg["ip"])
    a = (lambda k: None) if alt is None else (lambda k: alt[:k])
    co.solve(knot[:2048], theta[:2048], q2[:2048], o, threads=threads, alt=a(2048))  # warm-up (page-in, thread pool)
    t0 = time.perf_counter()
    z, dz, st, it = co.solve(knot[:sample], theta[:sample], q2[:sample], o, threads=threads, alt=a(sample))
    dt = time.perf_counter() - t0
    return sample / dt, co.max_threads if threads <= 0 else threads, dt, float(it.mean()), float(st.mean())


def config_block(cfg, world, rollouts, n):
    """The `config` object of the JSON line — identical in both arms (`--impl ours` / `--impl reference`)."""
    return {"workload": cfg["name"], "config_id": cfg["id"], "robot": cfg["robot"], "mode": cfg["mode"], "H_mpc": cfg["H"],
            "rollouts_per_gpu": int(rollouts), "subproblems_per_step_per_gpu": int(n), **cfg["ip"],
            "l2": "inputs+outputs per step exceed the 126 MB L2" if n * 3000 > 126e6 else "batch smaller than L2 (config size)",
            "parallelism": f"rollout-sharded x{world}, no data-path collective"}


def run_reference(args):
    """`--impl reference`: the reference's algorithm (C restatement: explicit Dx⁻¹, MGS-QR per
    iteration, all nθ sensitivity columns) on the host cores; rank 0 only.  Each step is a bounded sample (1/8 of the
    per-GPU batch, at most 81 920 subproblems) of the SAME workload as the `ours` arm; the metric is a rate."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = dict(CONFIGS[args.config], id=args.config)
    rollouts = args.rollouts or cfg["rollouts"]
    sample_rollouts = max(1, min(rollouts // 8, 8192)) if rollouts >= 64 else rollouts
    lin, knot, theta, q2, alt = build_workload(0, sample_rollouts, cfg)
    n = knot.shape[0]
    from common import SIZES
    from oracle.c_oracle import COracle
    from oracle.ip import IPOptions
    co = COracle(*SIZES[cfg["robot"]], lin, mode=cfg["mode"], solver="mgs")
    o = IPOptions(**cfg["ip"])
    for _ in range(max(args.warmup, 1)):
        co.solve(knot, theta, q2, o, alt=alt)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        z, dz, st, it = co.solve(knot, theta, q2, o, alt=alt)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    cores = co.max_threads
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_block(cfg, args.gpus, rollouts, rollouts * cfg["H"]),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} subproblems ({sample_rollouts} rollouts of the batch) x {args.steps} steps; C restatement "
                                   f"of the Julia CPU path (oracle/c/ip_oracle.c, OpenMP over problems)"},
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
    ap.add_argument("--config", type=int, default=4, choices=sorted(CONFIGS), help="BASELINE.json config (default 4)")
    ap.add_argument("--rollouts", type=int, default=None, help="rollouts per GPU (default: the config's)")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the extra.configs legs (configs 2, 3, 5)")
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

    numa = pin_to_gpu_numa(local)
    cfg = dict(CONFIGS[args.config], id=args.config)
    if args.rollouts is None:
        args.rollouts = cfg["rollouts"]
    ROBOT, MODE, H_MPC, IP_KW = cfg["robot"], cfg["mode"], cfg["H"], cfg["ip"]
    nq, nu, nw, nc, nb = SIZES[ROBOT]
    lin, knot, theta, q2, alt = build_workload(rank, args.rollouts, cfg)
    n = knot.shape[0]
    opts = cb.InteriorPointOptions(**IP_KW)
    im = cb.ImplicitTrajectory(nq, nu, nw, nc, nb, lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"],
                               mode=MODE, opts=opts, device=local)

    # ---- device-resident leg -------------------------------------------------------------
    kd = torch.from_numpy(knot).to(dev)
    td = torch.from_numpy(theta).to(dev)
    qd = torch.from_numpy(q2).to(dev)
    ad = torch.from_numpy(alt).to(dev) if alt is not None else None
    outb = im.solve_device(kd, td, qd, alt=ad)
    for _ in range(args.warmup):
        im.solve_device(kd, td, qd, alt=ad, out=outb)
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
        im.solve_device(kd, td, qd, alt=ad, out=outb)
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

    alt_h = torch.from_numpy(alt).pin_memory() if alt is not None else None
    if alt_h is not None:
        h2d += alt_h.numel() * 8

    def e2e_once(mask=None):
        im.solve_host_into(knot_h.numpy(), theta_h.numpy(), q2_h.numpy(), alt_h.numpy() if alt_h is not None else None,
                           z_h.numpy(), dz_h.numpy(), st_h.numpy(), it_h.numpy(), out_mask=mask)

    def timed_e2e(mask=None):
        e2e_once(mask)
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_once(mask)
        torch.cuda.synchronize()
        te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if dist:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item())

    e2e_s = timed_e2e()
    e2e_ok = bool(np.array_equal(z_h.numpy(), outb[0].cpu().numpy()))
    # the same call returning only what a caller of implicit_dynamics! that does not run Newton needs: d and the status
    mask_d = im.OUT_ZHEAD | im.OUT_STATUS
    e2e_d_s = timed_e2e(mask_d)
    d2h_d = n * im.nd * 8 + st_h.numel() + it_h.numel() * 4

    # ---- MPC-step leg: batched newton_solve! (cold start, one MPC step of every rollout) --------------
    mpc = None
    if not args.no_mpc and ROBOT == "quadruped":
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
        # warm-up: graph instantiation, lazy module loading, and the SM clocks coming back up after the bus-bound e2e leg
        # (the GPU idles there) — solve until two consecutive solves take the same time (±1.5 %), at most 16 times
        prev_ms = None
        for k in range(16):
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record()
            newton.solve(win, gait["q"][:H_MPC + 2], gait["u"][:H_MPC], gait["mu"], gait["h"], q0m, q1m)
            w1.record()
            torch.cuda.synchronize()
            ms_k = w0.elapsed_time(w1)
            if k >= 3 and prev_ms is not None and abs(ms_k - prev_ms) <= 0.015 * prev_ms:
                break
            prev_ms = ms_k
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        m_steps = 5
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
        # end to end at the path's own boundary: host q0, q1 in, host u out (cimpc_newton_solve_batch_host)
        q0h, q1h = q0m.cpu().pin_memory(), q1m.cpu().pin_memory()
        uh = torch.empty((Rm, nu), dtype=torch.float64).pin_memory()
        ih = torch.empty((Rm, 4), dtype=torch.int32).pin_memory()

        def mpc_host_once():
            newton.solve_host(win, gait["q"][:H_MPC + 2], gait["u"][:H_MPC], gait["mu"], gait["h"], q0h.numpy(), q1h.numpy(),
                              u_out=uh.numpy(), info_out=ih.numpy())
        mpc_host_once()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(m_steps):
            mpc_host_once()
        th = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if dist:
            dist.all_reduce(th, op=dist.ReduceOp.MAX)
        mpc_e2e = {"value": Rm * world * m_steps / float(th.item()), "unit": "MPC steps/s",
                   "h2d_bytes_per_step": int(2 * Rm * nq * 8), "d2h_bytes_per_step": int(Rm * nu * 8 + Rm * 16),
                   "matches_device_leg": bool(np.array_equal(uh.numpy(), um.cpu().numpy())),
                   "api": "cimpc_newton_solve_batch_host (pinned host q0, q1 in; host u, info out; one graph launch)"}
        mpc = {"value": Rm * world * m_steps / (float(tm.item()) * 1e-3), "unit": "MPC steps/s", "e2e": mpc_e2e,
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
    if not args.no_mpc and not args.no_closed_loop and ROBOT == "quadruped":
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
            # final gather of the (down-sampled: every N_sample-th step) configurations to rank 0: `cimpc_gather`, NCCL
            # point-to-point through the C ABI (torch.distributed only handed the 128-byte unique id round)
            cb.init_comm(im, world, rank)
            qg = cb.gather_rollouts_capi(im, res_c["q"].permute(1, 0, 2).contiguous(), Rc * world, world, rank)
            torch.cuda.synchronize()
            gathered = None if qg is None else list(qg.shape)
        else:
            gathered = list(res_c["q"].permute(1, 0, 2).shape)
        closed = {"value": Rc * world * mc.mpc_steps / (float(tc.item()) * 1e-3), "unit": "MPC steps/s (closed loop)",
                  "rollouts_per_gpu": Rc, "sim_steps": H_sim_c, "mpc_steps_per_rollout": mc.mpc_steps,
                  "ms_total": float(tc.item()), "sim_ok_frac": float(okc.item()), "groups": args.closed_loop_groups,
                  "gathered_trajectory_shape": gathered, "gather": "cimpc_gather (ncclSend/ncclRecv to rank 0)" if dist else "single rank",
                  "config": "examples/quadruped/monte_carlo.jl: initial configurations from the conf_min/conf_max box "
                            "(Philox seed 100), policy every 5 simulator steps, nonlinear simulator step on the device; "
                            "trajectories recorded every 5th step and gathered to rank 0"}

    # ---- the other BASELINE configs, so that the driver's record holds them (rank 0, one GPU each) ----------------
    extra = None
    if args.config == 4 and not args.no_extra_configs and rank == 0:
        extra = [dict(ip_sweep_leg(cb, torch, dev, dict(CONFIGS[c], id=c), 0, max(args.steps, 5), args.warmup), config_id=c)
                 for c in (2, 3, 5)]
    if dist:
        dist.barrier()
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
    fl_exe = executed_flops(nq, nu, nw, nc, nb, MODE, mean_it)
    f_peak, f_kind = fp64_peak()
    f_ach = n * fl / (kernel_ms * 1e-3) / 1e12
    f_exe = n * fl_exe / (kernel_ms * 1e-3) / 1e12
    # DRAM traffic of the kernel cannot be read without a profiler; the number below comes from the committed
    # `ncu --set full` capture of this round's kernel (per-subproblem streaming traffic × n), and says so.
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r02_ip_kernel_traffic.json")
    if os.path.exists(tp) and ROBOT == "quadruped":
        with open(tp) as f:
            tj = json.load(f)
        traffic = tj.get("dram_bytes_per_subproblem") * n
        traffic_src = f"profiles/r02_ip_kernel_traffic.json ({tj.get('source', 'ncu --set full')}), scaled to this launch; not measured in this run"

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_block(cfg, world, args.rollouts, n),
        "mean_ip_iterations": mean_it, "max_ip_iterations": max_it, "converged_frac": conv_frac,
        "clocks": clocks,
        "gpu_launches": int(launches),
        "e2e": {"value": n * world * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "matches_device_leg": e2e_ok,
                "api": "cimpc_ip_solve_batch_host (pinned host buffers, chunk-pipelined H2D / kernel / D2H; z, δz, status, "
                       "iters all returned)", "host_numa": numa,
                "d_and_status_only": {"value": n * world * e2e_steps / e2e_d_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                                      "d2h_bytes_per_step": int(d2h_d),
                                      "api": "cimpc_ip_solve_batch_host_ex, mask CIMPC_OUT_ZHEAD | CIMPC_OUT_STATUS (δz stays "
                                             "on the device: only newton_solve! consumes it)"}},
        "roofline": {"bound": "hbm", "kernel": "ip_solve_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": f"MEASURED_PEAKS.json ({peak_kind})", "algorithmic_bytes_per_subproblem": B,
                     "kernel_ms": kernel_ms,
                     "note": "fp64-issue bound path: HBM fraction is structurally small (SURVEY.md §8d); see fp64"},
        "fp64": {"peak": f_peak, "unit": "TFLOP/s", "peak_source": f_kind,
                 "executed": {"achieved": f_exe, "frac": f_exe / f_peak, "flops_per_subproblem": fl_exe,
                              "note": "flops this kernel executes (reduced 12x12 Gauss-Jordan inverse, 11 consumed sensitivity "
                                      "solves as products)"},
                 "reference_count": {"achieved": f_ach, "frac": f_ach / f_peak, "flops_per_subproblem": fl,
                                     "note": "flops the REFERENCE algorithm would spend (MGS-QR 16x16, all 34 sensitivity "
                                             "columns; BASELINE.md §3)"}},
    }
    if extra is not None:
        out["extra"] = {"configs": extra}
    if mpc is not None:
        out["mpc_steps"] = mpc
    if closed is not None:
        out["closed_loop"] = closed
    if not args.no_cpu_baseline and world == 1:
        sample = min(n, 262144)
        v, cores, dt, mit, conv = cpu_baseline_run(lin, knot, theta, q2, sample, cfg=cfg, alt=alt)
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
