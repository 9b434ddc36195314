"""Shared builders so a GPU field and its oracle twin are constructed from one description."""
import numpy as np

from opflow_b200 import capi, host
from oracle import oracle as O


def make_pair(dims, lo, hi, loc=0, bc=None, ext=0, padding=0, coords=None, name="u"):
    """-> (gpu Field, oracle Field) on identical meshes.  bc: {(axis, side): (type, value)}"""
    dim = len(dims)
    mb = host.MeshBuilder(dim).newMesh(*dims)
    for d in range(dim):
        if coords is not None and coords[d] is not None:
            mb.setMeshOfDim(d, np.asarray(coords[d], dtype=np.float64))
        else:
            mb.setMeshOfDim(d, lo[d], hi[d])
    gm = mb.build()
    om = O.Mesh(dims, lo, hi, coords=coords)
    return make_pair_on(gm, om, loc, bc, ext, padding, name)


def make_pair_on(gm, om, loc=0, bc=None, ext=0, padding=0, name="u"):
    dim = gm.dim
    b = host.ExprBuilder().setMesh(gm).setName(name).setLoc(loc)
    for (d, s), (t, v) in (bc or {}).items():
        b.setBC(d, s, t, v)
    if isinstance(ext, int):
        b.setExt(ext)
    else:
        for d in range(dim):
            for s in range(2):
                b.setExt(d, s, ext[d][s])
    b.setPadding(padding)
    g = b.build()
    o = O.Field(om, loc=loc, bc=bc, ext=ext, padding=padding, name=name)
    return g, o


def set_both(g, o, fn=None, arr=None):
    """same initial values into both twins over localRange"""
    if arr is None:
        rng = np.random.default_rng(1234)
        shape = [e - s for s, e in zip(*o.localRange)]
        arr = rng.standard_normal(shape) if fn is None else fn(shape)
    o.set_local(arr)
    g.from_numpy(arr)


def gpu_storage(g, o):
    """GPU values over the oracle's storage range ∩ readable range"""
    r = g.getLocalReadableRange()
    return g.to_numpy(r), o.view(r.tup(g.dim))


def assert_same(a, b, exact, tol=1e-12, what=""):
    if exact:
        assert np.array_equal(a, b), f"{what}: not bit-identical, max abs diff {np.abs(a - b).max()}"
    else:
        scale = max(np.abs(b).max(), 1e-300)
        err = np.abs(a - b).max() / scale
        assert err <= tol, f"{what}: relative L-inf error {err} > {tol}"


def dirc(dim, v=1.0):
    return {(d, s): (capi.BC_DIRC, v) for d in range(dim) for s in range(2)}
