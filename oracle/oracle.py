"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes/numpy view of oracle/liboracle.so (the plain-C restatement in oracle/opflow_oracle.c) plus helpers to run the
*real* reference drivers under oracle/_ref/bin (built by oracle/build_ref.sh from the unmodified reference sources).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the
product path (opflow_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "bin")

LOC_CORNER, LOC_CENTER = 0, 1
BC_UNDEFINED, BC_DIRC, BC_NEUM, BC_PERIODIC, BC_INTERNAL, BC_SYMM, BC_ASYMM = range(7)


class ORange(C.Structure):
    _fields_ = [("start", C.c_int * 3), ("end", C.c_int * 3)]

    def tup(self, dim=3):
        return tuple(self.start[:dim]), tuple(self.end[:dim])


class OField(C.Structure):
    _fields_ = [("dim", C.c_int), ("loc", C.c_int * 3), ("storage", ORange), ("local", ORange), ("assignable", ORange),
                ("accessible", ORange), ("logical", ORange), ("bc_type", (C.c_int * 2) * 3),
                ("bc_value", (C.c_double * 2) * 3), ("data", C.POINTER(C.c_double))]


class OMesh(C.Structure):
    _fields_ = [("dim", C.c_int), ("ext_start", C.c_int * 3), ("n_ext", C.c_int * 3),
                ("x", C.POINTER(C.c_double) * 3), ("dx", C.POINTER(C.c_double) * 3)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "opflow_oracle.c")):
            build()
        _lib = C.CDLL(LIB)
        _lib.orc_mesh_axis.restype = C.c_int
        _lib.orc_mesh_axis.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]
    return _lib


_dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))


class Mesh:
    """CartesianMesh restatement: per-axis x/dx/idx over the extended range (pad 5)."""

    def __init__(self, dims, lo=None, hi=None, coords=None, start=None, pad=5, ext_mode=None):
        self.dim = len(dims)
        self.dims = list(dims)
        self.start = list(start) if start else [0] * self.dim
        self.pad = pad
        self.x, self.dx, self.idx = [], [], []
        for d in range(self.dim):
            n = dims[d]
            x, dx, idx = np.zeros(n + 2 * pad), np.zeros(n + 2 * pad - 1), np.zeros(n + 2 * pad - 1)
            xs = None
            if coords is not None and coords[d] is not None:
                xs = np.ascontiguousarray(coords[d], dtype=np.float64)
            mode = 0 if ext_mode is None else ext_mode[d]
            lib().orc_mesh_axis(n, self.start[d], pad, mode, 0.0 if xs is not None else float(lo[d]),
                                0.0 if xs is not None else float(hi[d]), xs.ctypes.data if xs is not None else None,
                                x.ctypes.data, dx.ctypes.data, idx.ctypes.data)
            self.x.append(x), self.dx.append(dx), self.idx.append(idx)
        self.c = OMesh()
        self.c.dim = self.dim
        for d in range(self.dim):
            self.c.ext_start[d] = self.start[d] - pad
            self.c.n_ext[d] = dims[d] + 2 * pad
            self.c.x[d] = _dp(self.x[d])
            self.c.dx[d] = _dp(self.dx[d])

    def range(self):
        return self.start, [s + n for s, n in zip(self.start, self.dims)]


class Field:
    """CartesianField restatement with dense (unpitched) storage, exactly PlainTensor's layout."""

    def __init__(self, mesh: Mesh, loc=None, bc=None, ext=0, padding=0, name="", local_override=None):
        dim = mesh.dim
        self.mesh, self.name, self.dim = mesh, name, dim
        self.c = OField()
        self.c.dim = dim
        loc = [loc] * dim if isinstance(loc, int) else (loc or [LOC_CORNER] * dim)
        bc = bc or {}
        for d in range(dim):
            self.c.loc[d] = loc[d]
            for s in range(2):
                t, v = bc.get((d, s), (BC_UNDEFINED, 0.0))
                self.c.bc_type[d][s] = t
                self.c.bc_value[d][s] = v
        if isinstance(ext, int):
            ext = [[ext, ext]] * dim
        e = (C.c_int * (2 * dim))(*[w for p in ext for w in p])
        ms, me = mesh.range()
        lo = None
        if local_override is not None:
            lo = ORange()
            for d in range(3):
                lo.start[d] = local_override[0][d] if d < dim else 0
                lo.end[d] = local_override[1][d] if d < dim else 1
        self.padding = lib().orc_field_ranges(C.byref(self.c), (C.c_int * dim)(*ms), (C.c_int * dim)(*me), e, padding,
                                              C.byref(lo) if lo is not None else None)
        shape = [self.c.storage.end[d] - self.c.storage.start[d] for d in range(dim)]
        self.data = np.zeros(shape, order="F")
        self.c.data = _dp(self.data)
        self.update_padding()

    def _r(self, name):
        return getattr(self.c, name).tup(self.dim)

    localRange = property(lambda s: s._r("local"))
    assignableRange = property(lambda s: s._r("assignable"))
    accessibleRange = property(lambda s: s._r("accessible"))
    logicalRange = property(lambda s: s._r("logical"))
    storageRange = property(lambda s: s._r("storage"))

    def update_padding(self):
        lib().orc_update_padding(C.byref(self.c), C.byref(self.mesh.c))

    def view(self, rng):
        """numpy view of the values over rng = (start, end) global indices"""
        s0 = self.c.storage.start
        sl = tuple(slice(rng[0][d] - s0[d], rng[1][d] - s0[d]) for d in range(self.dim))
        return self.data[sl]

    def local(self):
        return self.view(self.localRange)

    def set_local(self, a):
        self.view(self.localRange)[...] = a
        self.update_padding()

    def writable(self):
        a, l = self.assignableRange, self.localRange
        return tuple(max(x, y) for x, y in zip(a[0], l[0])), tuple(min(x, y) for x, y in zip(a[1], l[1]))

    def init_by(self, f):
        """CartesianField::initBy (CartesianField.hpp:283-294)"""
        w = self.writable()
        v = self.view(w)
        m = self.mesh
        coords = []
        for d in range(self.dim):
            i = np.arange(w[0][d], w[1][d]) - (m.start[d] - m.pad)
            coords.append(m.x[d][i] if self.c.loc[d] == LOC_CORNER else m.x[d][i] + .5 * m.dx[d][i])
        it = np.nditer(v, flags=["multi_index"], op_flags=["writeonly"])
        for o in it:
            o[...] = f([coords[d][it.multi_index[d]] for d in range(self.dim)])
        self.update_padding()

    def fill(self, c):
        self.view(self.writable())[...] = c
        self.update_padding()


def _fields_arr(fields):
    return (C.POINTER(OField) * max(1, len(fields)))(*[C.pointer(f.c) for f in fields])


def assign(dst: Field, sig: str, fields, scalars=(), op=0):
    """dst (op)= expr, then updatePadding (FieldAssigner::assign + CartesianField::assignImpl_final)"""
    S = (C.c_double * max(1, len(scalars)))(*scalars)
    rc = lib().orc_assign(C.byref(dst.c), op, sig.encode(), _fields_arr(fields), S, C.byref(dst.mesh.c))
    if rc:
        raise RuntimeError(f"oracle: orc_assign({sig}) failed rc={rc}")


def evaluate(sig: str, fields, scalars, mesh: Mesh, lo, hi):
    dim = mesh.dim
    lo3 = list(lo) + [0] * (3 - dim)
    hi3 = list(hi) + [1] * (3 - dim)
    out = np.zeros([h - l for l, h in zip(lo3, hi3)][:dim], order="F")
    S = (C.c_double * max(1, len(scalars)))(*scalars)
    rc = lib().orc_eval(sig.encode(), _fields_arr(fields), S, C.byref(mesh.c), (C.c_int * 3)(*lo3), (C.c_int * 3)(*hi3), _dp(out))
    if rc:
        raise RuntimeError(f"oracle: orc_eval({sig}) failed rc={rc}")
    return out


def prepare(sig: str, fields, which=2):
    r, loc = ORange(), (C.c_int * 3)()
    rc = lib().orc_prepare(sig.encode(), _fields_arr(fields), which, C.byref(r), loc)
    if rc:
        raise RuntimeError(f"oracle: orc_prepare({sig}) failed rc={rc}")
    return r, list(loc)


def split_even(dim, mesh_start, mesh_end, nproc):
    out = (ORange * nproc)()
    lib().orc_split_even(dim, (C.c_int * dim)(*mesh_start), (C.c_int * dim)(*mesh_end), nproc, out)
    return [o.tup(dim) for o in out]


# ------------------------------------------------------------------------------------------- the real reference
def ref_available(name="ref_explicit"):
    return os.path.exists(os.path.join(REF_BIN, name))


def run_ref(name, *args, timeout=3600, env=None):
    """Run a reference driver from oracle/_ref/bin; returns the parsed JSON lines it prints."""
    exe = os.path.join(REF_BIN, name)
    out = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=timeout, env=env)
    if out.returncode != 0:
        raise RuntimeError(f"{name} failed ({out.returncode}): {out.stderr[-2000:]}")
    res = []
    for line in out.stdout.splitlines():
        line = line.strip()
        if line.startswith("{"):
            res.append(json.loads(line))
    return res


def read_opfd(path_or_bytes):
    """OPFD dump (oracle/ref_drivers/ref_common.hpp) -> (start, end, array[i0,i1,i2])"""
    b = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    out, off = [], 0
    while off < len(b):
        assert b[off:off + 4] == b"OPFD", "bad OPFD magic"
        dim = struct.unpack_from("i", b, off + 4)[0]
        se = struct.unpack_from(f"{2 * dim}i", b, off + 8)
        start, end = se[:dim], se[dim:]
        shape = [e - s for s, e in zip(start, end)]
        n = int(np.prod(shape))
        a = np.frombuffer(b, dtype="<f8", count=n, offset=off + 8 + 8 * dim).reshape(shape, order="F")
        out.append((tuple(start), tuple(end), a))
        off += 8 + 8 * dim + 8 * n
    return out[0] if len(out) == 1 else out
