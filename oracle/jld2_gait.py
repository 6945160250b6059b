"""Minimal read-only JLD2 (HDF5-subset) reader for the reference's gait files.

TEST INFRASTRUCTURE ONLY (part of `oracle/`): used once, in this container, to turn
the reference's gait fixtures (`/root/reference/src/dynamics/<robot>/gaits/*.jld2`,
loaded upstream by `get_trajectory(...; load_type = :split_traj_alt)`,
`src/controller/trajectory.jl:168-179`) into small `.npz` fixtures under
`tests/golden/` (see `oracle/make_golden.py`).  Nothing on the product path imports it.

Supported subset (what JLD2 0.1.1 writes for these files): v2 superblock at a 512 B
base offset, v2 object headers (OHDR/OCHK continuation), link messages, simple
dataspaces, fixed-point / floating-point / reference datatypes (incl. committed
datatypes), contiguous or compact layouts, no filters.
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


def load_split_traj_alt(path: str) -> dict:
    """Datasets read by `get_trajectory(..., load_type=:split_traj_alt)` (trajectory.jl:168-179)."""
    f = JLD2File(path)
    out = {}
    for k_jl, k in (("qm", "q"), ("um", "u"), ("γm", "gamma"), ("bm", "b"), ("ψm", "psi"), ("ηm", "eta")):
        out[k] = np.stack([np.asarray(v, dtype=np.float64) for v in f[k_jl]])
    out["mu"] = float(np.asarray(f["μm"]).reshape(-1)[0])
    out["h"] = float(np.asarray(f["hm"]).reshape(-1)[0])
    return out


def load_joint_traj(path: str) -> dict:
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
