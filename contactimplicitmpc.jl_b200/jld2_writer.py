"""JLD2 writer for the reference's gait files: `save_split_traj_alt` produces the file that
`@save path qm um γm bm ψm ηm μm hm` writes in the reference's trajectory-optimisation scripts (src/dynamics/utils.jl:129-152
and the generators under examples/) and that `get_trajectory(...; load_type = :split_traj_alt)` reads
(src/controller/trajectory.jl:168-179).

There is no Julia in this image, so the writer cannot be validated by JLD2.jl itself.  It is validated against JLD2.jl's
OUTPUT instead: re-encoding the contents of the reference's own gait files gives back the same bytes, checksums included
(tests/test_trajectory.py: every `:split_traj_alt` gait of the config robots when the reference tree is present, and
committed SHA-256 digests of three of them otherwise).  What is written is therefore exactly the layout JLD2 0.1.1
produced for these eight variables:

  offset   (relative to the 512-byte text header; all addresses in the file are relative to it as well)
  0        superblock v2 (lookup3 checksum)
  48       committed datatype `_types/00000001` — Julia's `DataType` as an HDF5 compound {name: vlen string,
           parameters: vlen of references}, attribute julia_type = itself
  200      global heap collection (4096 bytes): "Core.DataType", "Core.Array", "Core.Float64", the parameter list of
           Array{Float64,1} (two references)
  4312     three objects that spell the Julia type Array{Float64,1}: DataType(Core.Array, [↓, ↓]), DataType(Core.Float64, []),
           the Int64 1
  4527 …   for each of qm, um, γm, bm, ψm, ηm: a dataset of H object references (compact layout, attribute julia_type →
           4312) followed by its H element vectors, each a Float64 dataset with compact layout
  …        μm, hm (scalar Float64 datasets), the `_types` group, the root group (one link per variable, in write order)

Object headers are version 2 with a lookup3 (Bob Jenkins, `hashlittle`) checksum; sizes below 256 use a one-byte chunk
size, larger ones two bytes.
"""
from __future__ import annotations

import struct

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF
_BASE = 512


def _header_text(julia: str) -> bytes:
    return b"HDF5-based Julia Data Format, version 0.1.1\x00 (Julia " + julia.encode() + b" 64-bit LE)\x00"



def _rot(x: int, k: int) -> int:
    return ((x << k) | (x >> (32 - k))) & 0xFFFFFFFF


def lookup3(data: bytes, init: int = 0) -> int:
    """Bob Jenkins' lookup3 `hashlittle` — the checksum of HDF5 v2 superblocks and object headers."""
    m = 0xFFFFFFFF
    a = b = c = (0xDEADBEEF + len(data) + init) & m
    i, n = 0, len(data)
    while n > 12:
        a = (a + int.from_bytes(data[i:i + 4], "little")) & m
        b = (b + int.from_bytes(data[i + 4:i + 8], "little")) & m
        c = (c + int.from_bytes(data[i + 8:i + 12], "little")) & m
        a = (a - c) & m; a ^= _rot(c, 4); c = (c + b) & m
        b = (b - a) & m; b ^= _rot(a, 6); a = (a + c) & m
        c = (c - b) & m; c ^= _rot(b, 8); b = (b + a) & m
        a = (a - c) & m; a ^= _rot(c, 16); c = (c + b) & m
        b = (b - a) & m; b ^= _rot(a, 19); a = (a + c) & m
        c = (c - b) & m; c ^= _rot(b, 4); b = (b + a) & m
        i += 12
        n -= 12
    if n == 0:
        return c
    tail = data[i:] + b"\0" * (12 - n)
    a = (a + int.from_bytes(tail[0:4], "little")) & m
    b = (b + int.from_bytes(tail[4:8], "little")) & m
    c = (c + int.from_bytes(tail[8:12], "little")) & m
    c ^= b; c = (c - _rot(b, 14)) & m
    a ^= c; a = (a - _rot(c, 11)) & m
    b ^= a; b = (b - _rot(a, 25)) & m
    c ^= b; c = (c - _rot(b, 16)) & m
    a ^= c; a = (a - _rot(c, 4)) & m
    b ^= a; b = (b - _rot(a, 14)) & m
    c ^= b; c = (c - _rot(b, 24)) & m
    return c


# ---- header messages -----------------------------------------------------------------------------------------------
def _msg(mtype: int, payload: bytes, flags: int = 0) -> bytes:
    return struct.pack("<BHB", mtype, len(payload), flags) + payload


def _ohdr(messages: list[bytes]) -> bytes:
    """Version-2 object header, one chunk, no times / attribute phase-change fields."""
    body = b"".join(messages)
    if len(body) < 256:
        head = b"OHDR" + bytes([2, 0x00, len(body)])
    else:
        head = b"OHDR" + bytes([2, 0x01]) + struct.pack("<H", len(body))
    blob = head + body
    return blob + struct.pack("<I", lookup3(blob))


_FILL = _msg(0x05, bytes([3, 0x09]))                                   # fill value v3: never allocate, undefined
_SCALAR = _msg(0x01, bytes([2, 0, 0, 0]))                              # dataspace v2, scalar
_F64 = bytes([0x31, 0x20, 0x3F, 0x00]) + struct.pack("<IHHBBBBI", 8, 0, 64, 52, 11, 0, 52, 1023)  # IEEE double, LE
_I64 = bytes([0x30, 0x08, 0x00, 0x00]) + struct.pack("<IHH", 8, 0, 64)   # signed 64-bit integer, LE
_REF = bytes([0x37, 0x00, 0x00, 0x00]) + struct.pack("<I", 8)            # object reference
_NIL16 = _msg(0x00, bytes(16))                                          # the slack JLD2 leaves in group headers


def _simple(n: int) -> bytes:
    return _msg(0x01, bytes([2, 1, 0, 1]) + struct.pack("<Q", n))      # dataspace v2, rank 1


def _compact(data: bytes) -> bytes:
    return _msg(0x08, bytes([4, 0]) + struct.pack("<H", len(data)) + data)  # layout v4, compact


def _shared(addr: int) -> bytes:
    return bytes([3, 2]) + struct.pack("<Q", addr)                       # shared message v3 → committed datatype


def _attr_julia_type(datatype: bytes, shared: bool, data: bytes) -> bytes:
    name = b"julia_type\0"
    pl = bytes([2, 1 if shared else 0]) + struct.pack("<HHH", len(name), len(datatype), 4) + name + datatype + bytes([2, 0, 0, 0]) + data
    return _msg(0x0C, pl)


def _link(name: str, addr: int) -> bytes:
    nb = name.encode("utf-8")
    return _msg(0x06, bytes([1, 0x10, 1, len(nb)]) + nb + struct.pack("<Q", addr))  # v1, charset UTF-8, 1-byte length


def _group(links: list[tuple[str, int]], slack: int = 0) -> bytes:
    link_info = _msg(0x02, bytes([0, 0]) + struct.pack("<QQ", _UNDEF, _UNDEF))
    group_info = _msg(0x0A, bytes([0, 0]))
    extra = [_msg(0x00, bytes(slack))] if slack else []
    return _ohdr([link_info, group_info] + [_link(n, a) for n, a in links] + extra + [_NIL16])


def _vlen(length: int, heap: int, index: int) -> bytes:
    return struct.pack("<IQI", length, heap, index)


class _File:
    def __init__(self):
        self.buf = bytearray()

    def tell(self) -> int:
        return len(self.buf)

    def put(self, blob: bytes) -> int:
        at = len(self.buf)
        self.buf += blob
        return at

    def pad_to(self, addr: int):
        assert addr >= len(self.buf)
        self.buf += bytes(addr - len(self.buf))


def _type_preamble(f: _File) -> int:
    """Everything JLD2 writes before the first variable; returns the address of the object that spells Array{Float64,1}."""
    f.pad_to(48)  # the superblock goes in last
    # `DataType` as a committed compound datatype: {name: variable-length string; parameters: variable-length sequence of
    # object references}, 32 bytes
    vstr = bytes([0x39, 0x11, 0x01, 0x00]) + struct.pack("<I", 16) + bytes([0x30, 0x00, 0x00, 0x00]) + struct.pack("<IHH", 1, 0, 8)
    vref = bytes([0x39, 0x00, 0x00, 0x00]) + struct.pack("<I", 16) + _REF
    compound = bytes([0x36, 0x02, 0x00, 0x00]) + struct.pack("<I", 32) + b"name\0" + bytes([0]) + vstr + b"parameters\0" + bytes([16]) + vref
    heap = 200
    datatype_t = f.put(_ohdr([_msg(0x03, compound, 0x40),
                              _attr_julia_type(_shared(48), True, _vlen(13, heap, 1) + _vlen(0, 0, 0))]))
    assert datatype_t == 48
    # global heap collection
    f.pad_to(heap)
    array_t, float_t, one_t = 4312, 4391, 4470
    objs = [(1, b"Core.DataType"), (2, b"Core.Array"), (3, b"Core.Float64"), (4, struct.pack("<QQ", float_t, one_t))]
    col = b"GCOL" + bytes([1, 0, 0, 0]) + struct.pack("<Q", 4096)
    for idx, data in objs:
        col += struct.pack("<HHIQ", idx, 1, 0, len(data)) + data + bytes(-len(data) % 8)
    col += struct.pack("<HHIQ", 0, 0, 0, 4096 - len(col))  # the free-space object spans the rest of the collection
    f.put(col)
    f.pad_to(array_t)
    inst = lambda name_len, name_idx, n_par, par_idx: _ohdr([  # noqa: E731  an instance of `DataType`
        _FILL, _SCALAR, _msg(0x03, _shared(48), 0x03),
        _compact(_vlen(name_len, heap, name_idx) + (_vlen(n_par, heap, par_idx) if n_par else _vlen(0, 0, 0)))])
    assert f.put(inst(10, 2, 2, 4)) == array_t        # Core.Array{…}: parameters = heap object 4
    assert f.put(inst(12, 3, 0, 0)) == float_t        # Core.Float64
    assert f.put(_ohdr([_FILL, _SCALAR, _msg(0x03, _I64, 0x01), _compact(struct.pack("<q", 1))])) == one_t  # N = 1
    return array_t


def _float_vector(v: np.ndarray) -> bytes:
    data = np.ascontiguousarray(v, dtype="<f8").tobytes()
    return _ohdr([_FILL, _simple(v.size), _msg(0x03, _F64, 0x01), _compact(data)])


def _float_scalar(x: float) -> bytes:
    return _ohdr([_FILL, _SCALAR, _msg(0x03, _F64, 0x01), _compact(struct.pack("<d", float(x)))])


def encode_split_traj_alt(q, u, gamma, b, psi, eta, mu: float, h: float, julia: str = "1.6.0") -> bytes:
    """The bytes of a `:split_traj_alt` gait file: q (H+2, nq); u, γ, b, ψ, η (H, ·); μ, h scalars.
    `julia`: the Julia version named in the text header.  The reference's quadruped / flamingo / hopper gaits were written
    under 1.6.0; its centroidal gaits under 1.7.2 and 1.8.2, whose JLD2 release leaves 68 more bytes of slack (a NIL
    message) in the `_types` group header — the only difference in the bytes."""
    f = _File()
    array_t = _type_preamble(f)
    links = []
    for name, arr in (("qm", q), ("um", u), ("γm", gamma), ("bm", b), ("ψm", psi), ("ηm", eta)):
        arr = np.asarray(arr, dtype=np.float64)
        assert arr.ndim == 2
        elems = [_float_vector(row) for row in arr]
        # the header of the outer vector comes first; its size is known without the addresses
        probe = _ohdr([_FILL, _simple(len(elems)), _attr_julia_type(_REF, False, struct.pack("<Q", array_t)),
                       _msg(0x03, _REF, 0x01), _compact(bytes(8 * len(elems)))])
        at = f.tell()
        addr = at + len(probe)
        refs = []
        for e in elems:
            refs.append(addr)
            addr += len(e)
        f.put(_ohdr([_FILL, _simple(len(elems)), _attr_julia_type(_REF, False, struct.pack("<Q", array_t)),
                     _msg(0x03, _REF, 0x01), _compact(struct.pack(f"<{len(refs)}Q", *refs))]))
        for e in elems:
            f.put(e)
        links.append((name, at))
    links.append(("μm", f.put(_float_scalar(mu))))
    links.append(("hm", f.put(_float_scalar(h))))
    newer = tuple(int(x) for x in julia.split(".")[:2]) >= (1, 7)
    types = f.put(_group([("00000001", 48)], slack=68 if newer else 0))
    links.append(("_types", types))
    root = f.put(_group(links))
    eof = _BASE + f.tell()
    sb = b"\x89HDF\r\n\x1a\n" + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", _BASE, _UNDEF, eof, root)
    f.buf[0:48] = sb + struct.pack("<I", lookup3(sb))
    text = _header_text(julia)
    return text + bytes(_BASE - len(text)) + bytes(f.buf)


def save_split_traj_alt(path: str, q, u, gamma, b, psi, eta, mu: float, h: float, julia: str = "1.6.0"):
    """`@save path qm um γm bm ψm ηm μm hm`: a gait file `get_trajectory(...; load_type = :split_traj_alt)` reads."""
    with open(path, "wb") as fh:
        fh.write(encode_split_traj_alt(q, u, gamma, b, psi, eta, mu, h, julia))
