"""Reference-trajectory handling of the host mirror: the gait files and `ContactTraj`.

  * `JLD2File`, `load_gait`     read the reference's gait fixtures (`src/dynamics/<robot>/gaits/*.jld2`) as
                                `get_trajectory(model, env, path; load_type)` does (src/controller/trajectory.jl:143-185):
                                `:split_traj_alt` = datasets qm, um, γm, bm, ψm, ηm, μm, hm (:168-179),
                                `:joint_traj` = one serialized `ContactTraj` under "traj" (:180-181)
  * `ContactTraj`               q, u, w, γ, b, z, θ of a reference or simulated trajectory (trajectory.jl:1-82),
                                `update_z!`, `update_θ!`, `pack_z` (src/simulation/index.jl:437-441)
  * `repeat_ref_traj`, `tracking_error`   trajectory.jl:84-113, 188-217 — the closed-loop metrics the reference's
                                end-to-end tests pin (test/controller/mpc_quadruped.jl:59-64)
  * `save_traj` / `load_traj`   npz round trip of a whole `ContactTraj`
  * `save_gait`                 writes a `:split_traj_alt` JLD2 gait file (`jld2_writer.py`: byte-for-byte what JLD2.jl
                                wrote for the reference's own gaits; the reference uses `@save`, src/dynamics/utils.jl:129-152)

A Julia caller keeps using JLD2.jl; this reader exists so that the Python mirror consumes the SAME files.
Supported HDF5 subset (what JLD2 0.1.1 wrote for these files): v2 superblock at a 512-byte base offset, v2 object
headers with continuation blocks, link messages, simple dataspaces, fixed-point / floating-point / reference
datatypes (incl. committed ones), contiguous or compact layouts, no filters.
"""
from __future__ import annotations

import struct

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF


class JLD2File:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.buf = f.read()
        sig = b"\x89HDF\r\n\x1a\n"
        self.base = self.buf.find(sig)
        if self.base < 0:
            raise ValueError("not an HDF5/JLD2 file")
        b = self.base
        ver = self.buf[b + 8]
        if ver not in (2, 3):
            raise ValueError(f"unsupported superblock version {ver}")
        if self.buf[b + 9] != 8 or self.buf[b + 10] != 8:
            raise ValueError("only 8-byte offsets/lengths supported")
        self.base_addr, _ext, _eof, self.root = struct.unpack_from("<QQQQ", self.buf, b + 12)
        self.links = self._links(self.root)

    # -- object headers -------------------------------------------------
    def _messages(self, addr: int):
        """Yield (type, payload bytes) for every message of the v2 object header at addr."""
        buf = self.buf
        p = self.base_addr + addr
        if buf[p:p + 4] != b"OHDR":
            raise ValueError(f"no OHDR at {addr:#x}")
        if buf[p + 4] != 2:
            raise ValueError("only v2 object headers supported")
        flags = buf[p + 5]
        p += 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        nsz = 1 << (flags & 0x3)
        chunk = int.from_bytes(buf[p:p + nsz], "little")
        p += nsz
        blocks = [(p, p + chunk)]
        out = []
        while blocks:
            p, end = blocks.pop(0)
            while p + 4 <= end:
                mtype = buf[p]
                msize = struct.unpack_from("<H", buf, p + 1)[0]
                p += 4
                if flags & 0x04:
                    p += 2
                payload = buf[p:p + msize]
                p += msize
                if mtype == 0x10:  # continuation
                    off, length = struct.unpack_from("<QQ", payload, 0)
                    q = self.base_addr + off
                    if buf[q:q + 4] != b"OCHK":
                        raise ValueError("bad continuation block")
                    blocks.append((q + 4, q + length - 4))
                elif mtype != 0:
                    out.append((mtype, payload))
        return out

    def _links(self, addr: int) -> dict:
        links = {}
        for mtype, pl in self._messages(addr):
            if mtype != 6:
                continue
            flags = pl[1]
            p = 2
            ltype = 0
            if flags & 0x08:
                ltype = pl[p]
                p += 1
            if flags & 0x04:
                p += 8
            if flags & 0x10:
                p += 1
            nsz = 1 << (flags & 0x3)
            nlen = int.from_bytes(pl[p:p + nsz], "little")
            p += nsz
            name = pl[p:p + nlen].decode("utf-8")
            p += nlen
            if ltype == 0:
                links[name] = struct.unpack_from("<Q", pl, p)[0]
        return links

    # -- datasets ------------------------------------------------------
    def _datatype(self, pl: bytes):
        cls = pl[0] & 0x0F
        size = struct.unpack_from("<I", pl, 4)[0]
        return cls, size

    def read(self, addr: int):
        shape = None
        dtype = None
        data = None
        for mtype, pl in self._messages(addr):
            if mtype == 1:  # dataspace
                ver, ndim, fl = pl[0], pl[1], pl[2]
                p = 8 if ver == 1 else 4
                shape = struct.unpack_from("<" + "Q" * ndim, pl, p) if ndim else ()
            elif mtype == 3:  # datatype
                ver = pl[0] >> 4
                if ver == 0 and False:
                    pass
                dtype = self._datatype(pl)
            elif mtype == 8:  # layout
                ver, lcls = pl[0], pl[1]
                if ver not in (3, 4):
                    raise ValueError("layout version")
                if lcls == 1:
                    a, n = struct.unpack_from("<QQ", pl, 2)
                    data = None if a == _UNDEF else self.buf[self.base_addr + a:self.base_addr + a + n]
                elif lcls == 0:
                    n = struct.unpack_from("<H", pl, 2)[0]
                    data = pl[4:4 + n]
                else:
                    raise ValueError("chunked layout unsupported")
        if dtype is None:
            # shared (committed) datatype message: flag bit 1 of the message; find it raw
            dtype = self._shared_datatype(addr)
        count = int(np.prod(shape)) if shape is not None and len(shape) else 1
        cls, size = dtype
        if data is None:
            data = b""
        if cls == 1 and size == 8:
            arr = np.frombuffer(data, dtype="<f8", count=count).copy()
        elif cls == 0 and size == 8:
            arr = np.frombuffer(data, dtype="<i8", count=count).copy()
        elif cls == 7:  # object references
            refs = np.frombuffer(data, dtype="<u8", count=count)
            return [self.read(int(r)) for r in refs]
        else:
            raise ValueError(f"unsupported datatype class {cls} size {size}")
        if shape is not None and len(shape) > 1:
            # HDF5 dims are C-order of the reversed Julia dims
            arr = arr.reshape(shape).T
        elif shape == ():
            return arr[0]
        return arr

    def _shared_datatype(self, addr: int):
        # re-walk messages keeping flags: committed datatypes appear as type 3 with
        # the 'shared' flag; payload = version(1) type(1) address(8)
        buf = self.buf
        for mtype, pl in self._messages_with_flags(addr):
            if mtype[0] == 3 and (mtype[1] & 0x02):
                a = struct.unpack_from("<Q", pl, 2)[0]
                for t2, pl2 in self._messages(a):
                    if t2 == 3:
                        return self._datatype(pl2)
        raise ValueError("datatype not found")

    def _messages_with_flags(self, addr: int):
        buf = self.buf
        p = self.base_addr + addr
        flags = buf[p + 5]
        p += 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        nsz = 1 << (flags & 0x3)
        chunk = int.from_bytes(buf[p:p + nsz], "little")
        p += nsz
        blocks = [(p, p + chunk)]
        out = []
        while blocks:
            p, end = blocks.pop(0)
            while p + 4 <= end:
                mtype = buf[p]
                msize = struct.unpack_from("<H", buf, p + 1)[0]
                mflags = buf[p + 3]
                p += 4
                if flags & 0x04:
                    p += 2
                payload = buf[p:p + msize]
                p += msize
                if mtype == 0x10:
                    off, length = struct.unpack_from("<QQ", payload, 0)
                    q = self.base_addr + off
                    blocks.append((q + 4, q + length - 4))
                elif mtype != 0:
                    out.append(((mtype, mflags), payload))
        return out

    def __getitem__(self, name: str):
        return self.read(self.links[name])

    def keys(self):
        return list(self.links.keys())


def _load_split_traj_alt(path: str) -> dict:
    """Datasets read by `get_trajectory(..., load_type=:split_traj_alt)` (trajectory.jl:168-179)."""
    f = JLD2File(path)
    out = {}
    for k_jl, k in (("qm", "q"), ("um", "u"), ("γm", "gamma"), ("bm", "b"), ("ψm", "psi"), ("ηm", "eta")):
        out[k] = np.stack([np.asarray(v, dtype=np.float64) for v in f[k_jl]])
    out["mu"] = float(np.asarray(f["μm"]).reshape(-1)[0])
    out["h"] = float(np.asarray(f["hm"]).reshape(-1)[0])
    return out


def _load_joint_traj(path: str) -> dict:
    """`get_trajectory(..., load_type = :joint_traj)` (trajectory.jl:180-181): the file holds ONE serialized
    `ContactTraj` struct under "traj" (scalar dataset, committed compound datatype, compact layout):
    H::Int64, h::Float64, then 15 object references κ, q, u, w, γ, b, z, θ, iq0, iq1, iu1, iw1, iq2, iγ1, ib1
    (field order of `struct ContactTraj`, trajectory.jl:1-19).  Returns the arrays needed downstream."""
    f = JLD2File(path)
    payload = None
    for (mtype, _fl), pl in f._messages_with_flags(f.links["traj"]):
        if mtype == 8:
            if pl[1] != 0:
                raise ValueError("expected a compact layout for the ContactTraj struct")
            n = struct.unpack_from("<H", pl, 2)[0]
            payload = pl[4:4 + n]
    if payload is None or len(payload) != 16 + 15 * 8:
        raise ValueError("unexpected ContactTraj layout")
    H, = struct.unpack_from("<q", payload, 0)
    h, = struct.unpack_from("<d", payload, 8)
    refs = struct.unpack_from("<15Q", payload, 16)
    out = {"H": int(H), "h": float(h)}
    for name, ref in zip(("kappa", "q", "u", "w", "gamma", "b", "z", "theta"), refs[:8]):
        v = f.read(int(ref))
        out[name] = np.stack([np.asarray(x, dtype=np.float64) for x in v]) if isinstance(v, list) else np.asarray(v)
    assert out["q"].shape[0] == H + 2 and out["theta"].shape[0] == H
    return out


def load_gait(path: str, load_type: str = "split_traj_alt") -> dict:
    """`get_trajectory(...; load_type)`: dict with q (H+2, nq), u, gamma, b (H, ·), h and — split files — psi, eta, mu;
    — joint files — z, theta, w, kappa, H."""
    if load_type == "split_traj_alt":
        return _load_split_traj_alt(path)
    if load_type == "joint_traj":
        return _load_joint_traj(path)
    raise ValueError(f"unknown load_type {load_type!r}")


# ------------------------------------------------------------------------------------------------------
_PHI = {}


def _phi(robot: str):
    """ϕ(q) of the robot as a numpy function (from the same sympy model the CUDA code is generated from)."""
    if robot not in _PHI:
        import sympy as sp
        from .modelgen.robots import get_model
        m = get_model(robot)
        qs = sp.symbols(f"q0:{m.nq}", real=True)
        f = sp.lambdify([qs], [sp.sympify(e) for e in m.phi_func(qs)], "numpy")
        _PHI[robot] = (m, lambda q: np.array(f(list(q)), dtype=np.float64))
    return _PHI[robot]


class ContactTraj:
    """trajectory.jl:1-49.  q: (H+2, nq); u, w, γ, b, z, θ: (H, ·); z = [q2; γ1; b1; ψ1; s1; η1; s2], θ = [q0; q1; u1; w1; μ; h]."""

    def __init__(self, nq, nu, nw, nc, nb, H, h, kappa=0.0):
        self.nq, self.nu, self.nw, self.nc, self.nb = nq, nu, nw, nc, nb
        self.H, self.h, self.kappa = int(H), float(h), float(kappa)
        self.nz, self.ntheta = nq + 4 * nc + 2 * nb, 2 * nq + nu + nw + 2
        self.q = np.zeros((H + 2, nq)); self.u = np.zeros((H, nu)); self.w = np.zeros((H, nw))
        self.gamma = np.zeros((H, nc)); self.b = np.zeros((H, nb))
        self.z = np.zeros((H, self.nz)); self.theta = np.zeros((H, self.ntheta))
        self.theta[:, -1] = h  # trajectory.jl:38

    def update_z(self):  # update_z!, trajectory.jl:51-65
        nq, nc, nb = self.nq, self.nc, self.nb
        self.z[:, :nq] = self.q[2:]
        self.z[:, nq:nq + nc] = self.gamma
        self.z[:, nq + nc:nq + nc + nb] = self.b

    def update_theta(self):  # update_θ!, trajectory.jl:67-82
        nq, nu, nw = self.nq, self.nu, self.nw
        self.theta[:, :nq] = self.q[:-2]
        self.theta[:, nq:2 * nq] = self.q[1:-1]
        self.theta[:, 2 * nq:2 * nq + nu] = self.u
        self.theta[:, 2 * nq + nu:2 * nq + nu + nw] = self.w

    @classmethod
    def from_gait(cls, robot: str, gait: dict, kappa: float = 0.0) -> "ContactTraj":
        """`get_trajectory` tail (trajectory.jl:168-185): builds z through `pack_z` (s1 = ϕ(q2), s2 = μ_world γ1 − Σ b1;
        index.jl:437-441 — note `model.μ_world`, not the gait's μ) and θ with the gait's μ, h."""
        m, phi = _phi(robot)
        if "z" in gait and "theta" in gait:  # :joint_traj
            H = gait["u"].shape[0]
            tr = cls(m.nq, m.nu, m.nw, m.nc, m.nb, H, float(gait["h"]), kappa)
            for k in ("q", "u", "w", "gamma", "b", "z", "theta"):
                getattr(tr, k)[:] = gait[k]
            return tr
        H = gait["u"].shape[0]
        tr = cls(m.nq, m.nu, m.nw, m.nc, m.nb, H, float(gait["h"]), kappa)
        tr.q[:], tr.u[:], tr.gamma[:], tr.b[:] = gait["q"], gait["u"], gait["gamma"], gait["b"]
        nf = m.nb // m.nc
        for t in range(H):
            s1 = phi(gait["q"][t + 2])
            s2 = m.mu_world * gait["gamma"][t] - gait["b"][t].reshape(m.nc, nf).sum(axis=1)
            tr.z[t] = np.concatenate([gait["q"][t + 2], gait["gamma"][t], gait["b"][t], gait["psi"][t], s1, gait["eta"][t], s2])
        tr.update_theta()
        tr.theta[:, -2] = float(gait["mu"])
        return tr


def save_traj(path: str, tr: ContactTraj):
    np.savez_compressed(path, q=tr.q, u=tr.u, w=tr.w, gamma=tr.gamma, b=tr.b, z=tr.z, theta=tr.theta, h=tr.h,
                        kappa=tr.kappa, sizes=np.array([tr.nq, tr.nu, tr.nw, tr.nc, tr.nb]))


def save_gait(path: str, tr: ContactTraj, mu: float | None = None, julia: str = "1.6.0"):
    """`@save path qm um γm bm ψm ηm μm hm` of a trajectory: the file `get_trajectory(...; load_type = :split_traj_alt)`
    reads.  ψ and η are taken from z (index.jl: z = [q2; γ1; b1; ψ1; s1; η1; s2]); μ defaults to the one stored in θ."""
    from .jld2_writer import save_split_traj_alt
    nq, nc, nb = tr.nq, tr.nc, tr.nb
    psi = tr.z[:, nq + nc + nb:nq + 2 * nc + nb]
    eta = tr.z[:, nq + 3 * nc + nb:nq + 3 * nc + 2 * nb]
    save_split_traj_alt(path, tr.q, tr.u, tr.gamma, tr.b, psi, eta, float(tr.theta[0, -2]) if mu is None else mu, tr.h, julia)


def load_traj(path: str) -> ContactTraj:
    with np.load(path) as f:
        nq, nu, nw, nc, nb = (int(v) for v in f["sizes"])
        tr = ContactTraj(nq, nu, nw, nc, nb, f["u"].shape[0], float(f["h"]), float(f["kappa"]))
        for k in ("q", "u", "w", "gamma", "b", "z", "theta"):
            getattr(tr, k)[:] = f[k]
    return tr


def repeat_ref_traj(ref: ContactTraj, N: int, idx_shift=(0,)):
    """trajectory.jl:84-113: the reference repeated N times, each copy shifted by the stride along `idx_shift`."""
    idx = list(idx_shift)
    shift = (ref.q[-1] - ref.q[1])[idx]
    q, u, g, b = [ref.q.copy()], [ref.u], [ref.gamma], [ref.b]
    for k in range(1, N):
        qt = ref.q[2:].copy()
        qt[:, idx] += k * shift
        q.append(qt); u.append(ref.u); g.append(ref.gamma); b.append(ref.b)
    return np.concatenate(q), np.concatenate(u), np.concatenate(g), np.concatenate(b)


def tracking_error(ref: ContactTraj, sim_q, sim_u, sim_gamma, sim_b, N_sample: int, idx_shift=(0,)) -> np.ndarray:
    """trajectory.jl:188-217: mean ℓ1 errors (q, u, γ, b) of a simulated trajectory against the repeated reference at
    the MPC sampling instants.  sim_q: (H_sim+2, nq); sim_u, sim_gamma, sim_b: (H_sim, ·) at the simulator rate
    (u compared as recorded, i.e. the per-simulator-step impulse)."""
    H_sim = sim_u.shape[0]
    n_rep = int(np.ceil(H_sim / N_sample / ref.H))
    dq, du, dg, db = repeat_ref_traj(ref, n_rep, idx_shift)
    e = np.zeros(4)
    cnt = 0
    for t in range(1, ref.H * n_rep + 1):  # the reference's 1-based loop, including its count-before-break quirk
        cnt += 1
        if (t - 1) * N_sample + 1 > H_sim:
            break
        k = (t - 1) * N_sample
        e[0] += np.abs(dq[t + 1] - sim_q[k + 2]).sum() / ref.nq
        e[1] += np.abs(du[t - 1] - sim_u[k]).sum() / ref.nu
        e[2] += np.abs(dg[t - 1] - sim_gamma[k]).sum() / ref.nc
        e[3] += np.abs(db[t - 1] - sim_b[k]).sum() / ref.nb
    return e / cnt
