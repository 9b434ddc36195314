"""opf_assign_host (host buffers, upload | sweep | download pipelined in slabs) against the resident path on identical data --
bit for bit, both arithmetic modes, aliased and non-aliased destinations, 2-D and 3-D, plus the sequential fallback
(fields with ghost cells)."""
import ctypes as C

import numpy as np
import pytest

from opflow_b200 import capi, host
from opflow_b200.host import D2SecondOrderCentered as D2, d2x, d2y, d2z

pytestmark = pytest.mark.gpu


def field(dims, ext=0, name="u"):
    dim = len(dims)
    mb = host.MeshBuilder(dim).newMesh(*dims)
    for d in range(dim):
        mb.setMeshOfDim(d, 0., 1. + 0.5 * d)
    b = host.ExprBuilder().setName(name).setMesh(mb.build()).setExt(ext)
    for d in range(dim):
        b.setBC(d, 0, host.BCType.Dirc, 1.).setBC(d, 1, host.BCType.Dirc, 0.25)
    return b.build()


def lap(u, dim):
    e = d2x(D2, u) + d2y(D2, u)
    return e + d2z(D2, u) if dim == 3 else e


@pytest.mark.parametrize("dims,ext", [((70, 37, 49), 0), ((141, 130), 0), ((33, 18, 40), 1)])
@pytest.mark.parametrize("mode", [capi.MODE_EXACT, capi.MODE_FAST])
def test_host_assign_equals_resident_assign(engine, dims, ext, mode):
    l = engine
    host.set_mode(mode)
    dim = len(dims)
    rng = np.random.default_rng(3)
    u, r = field(dims, ext, "u"), field(dims, ext, "r")
    lr = u.localRange
    c = 0.02 * min((1. + 0.5 * d) / (dims[d] - 1) for d in range(dim)) ** 2
    steps = 3
    inits = [np.asfortranarray(rng.standard_normal(lr.shape(dim))) for _ in range(steps)]
    # resident reference: upload, updatePadding (the uploaded values ARE the field: its BC ghosts follow them), assign, download
    want = []
    for a in inits:
        r.upload_raw(a.ctypes.data, lr)
        r.updatePadding()
        r.assign(r + c * lap(r, dim))
        want.append(r.to_numpy())
    sig, fields, scalars = (u + c * lap(u, dim)).flatten()
    F = (C.c_void_p * len(fields))(*[f.h for f in fields])
    S = (C.c_double * len(scalars))(*scalars)
    for a, w in zip(inits, want):
        out = np.empty(lr.shape(dim), order="F")
        capi.check(l.opf_assign_host(u.h, capi.OP_EQ, sig.encode(), F, len(fields), S, len(scalars), u.h, C.c_void_p(a.ctypes.data),
                                     C.c_void_p(out.ctypes.data)))
        assert np.array_equal(out, w), np.abs(out - w).max()
        assert np.array_equal(u.to_numpy(), w)


def test_host_assign_other_destination(engine):
    """v = lap(u) with u streamed from the host: non-aliased destination"""
    l = engine
    host.set_mode(capi.MODE_EXACT)
    dims = (40, 33, 64)
    u, v, ur, vr = field(dims, 0, "u"), field(dims, 0, "v"), field(dims, 0, "ur"), field(dims, 0, "vr")
    lr = u.localRange
    a = np.asfortranarray(np.random.default_rng(5).standard_normal(lr.shape(3)))
    ur.upload_raw(a.ctypes.data, lr)
    ur.updatePadding()
    vr.assign(lap(ur, 3))
    sig, fields, scalars = lap(u, 3).flatten()
    F = (C.c_void_p * len(fields))(*[f.h for f in fields])
    out = np.empty(lr.shape(3), order="F")
    capi.check(l.opf_assign_host(v.h, capi.OP_EQ, sig.encode(), F, len(fields), None, 0, u.h, C.c_void_p(a.ctypes.data), C.c_void_p(out.ctypes.data)))
    assert np.array_equal(out, vr.to_numpy())
