"""The CUDA path through the C ABI against fixtures produced by the UNMODIFIED reference (tests/golden/, generator
oracle/make_golden.py): index maps / boundary classification / ghost values / explicit fields, all bit-exact in EXACT mode."""
import json
import os

import numpy as np
import pytest

from opflow_b200 import capi, host
from opflow_b200.host import (D1FirstOrderBiasedDownwind, D1WENO53Downwind, D1WENO53Upwind, D2SecondOrderCentered as D2, d2x, d2y, d2z,
                              dx)
from oracle import oracle as O
from test_oracle_pinned import BC, GOLD, REF, stretched

pytestmark = pytest.mark.gpu


def build(dims, lo, hi, coords, loc, bc, ext):
    dim = len(dims)
    mb = host.MeshBuilder(dim).newMesh(*dims)
    for d in range(dim):
        mb.setMeshOfDim(d, coords[d]) if coords is not None else mb.setMeshOfDim(d, lo[d], hi[d])
    b = host.ExprBuilder().setMesh(mb.build()).setLoc(loc).setExt(ext)
    for (d, s), (t, v) in bc.items():
        if t != capi.BC_UNDEFINED:
            b.setBC(d, s, t, v)
    return b.build()


def rng(r, dim):
    return [list(r.tup(dim)[0]), list(r.tup(dim)[1])]


def test_ranges_and_ghosts_1d(engine):
    host.set_mode(capi.MODE_EXACT)
    f = lambda x: 1.0 + 0.5 * x[0] + x[0] * x[0]
    for c in REF["ranges1d"]:
        n = c["n"]
        bc = {(0, 0): (BC[c["bc"][0]], c["bcv"][0]), (0, 1): (BC[c["bc"][1]], c["bcv"][1])}
        u = build([n], [0.0], [2.0], [stretched(n)] if c["stretched"] else None, c["loc"], bc, c["ext"])
        r = c["ranges"]
        assert rng(u.localRange, 1) == r["local"] and rng(u.assignableRange, 1) == r["assignable"], c
        assert rng(u.accessibleRange, 1) == r["accessible"] and rng(u.logicalRange, 1) == r["logical"] and u.padding == r["padding"], c
        u.initBy(f)
        fr = c["field"]["range"]
        got = u.to_numpy(capi.Range.make(fr[0], fr[1]))
        ref = np.array(c["field"]["values"])
        lo = r["local"][0][0] if c["bc"][0] == "Undefined" else fr[0][0]
        hi = r["local"][1][0] if c["bc"][1] == "Undefined" else fr[1][0]
        sl = slice(lo - fr[0][0], hi - fr[0][0])
        assert np.array_equal(got[sl], ref[sl]), c


def test_ghosts_2d_axis_order(engine):
    f = lambda x: np.sin(1.3 * x[0]) + x[1] * x[1] + 0.25 * x[0] * x[1]
    for c in REF["ghost2d"]:
        nx, ny = c["dims"]
        bc = {}
        for d in range(2):
            t0, v0, t1, v1 = c["bc"][d]
            bc[(d, 0)], bc[(d, 1)] = (BC[t0], float(v0)), (BC[t1], float(v1))
        u = build([nx, ny], [0.0, 0.0], [2.0, 1.0], [stretched(nx, 2.0), stretched(ny)] if c["stretched"] else None, c["loc"], bc, c["ext"])
        u.initBy(f)
        fr = c["field"]["range"]
        ref = np.array(c["field"]["values"]).reshape([fr[1][0] - fr[0][0], fr[1][1] - fr[0][1]], order="F")
        got = u.to_numpy(capi.Range.make(fr[0], fr[1]))
        assert rng(u.logicalRange, 2) == c["ranges"]["logical"]
        assert np.array_equal(got, ref), (c["loc"], c["bc"], c["stretched"], np.argwhere(got != ref)[:4])


def test_prepared_expression_ranges(engine):
    bc = {(0, 0): (capi.BC_DIRC, 1.0), (0, 1): (capi.BC_NEUM, 0.0), (1, 0): (capi.BC_NEUM, 0.0), (1, 1): (capi.BC_DIRC, 0.0)}
    import ctypes as C
    l = capi.lib()
    for c in REF["prepare2d"]:
        u = build([12, 10], [0.0, 0.0], [2.0, 1.0], None, c["loc"], bc, 3)
        nf = c["sig"].count("F<")
        F = (C.c_void_p * nf)(*[u.h] * nf)
        for which, key in ((capi.R_ACCESSIBLE, "acc"), (capi.R_LOCAL, "local"), (capi.R_LOGICAL, "logical")):
            r, loc = capi.Range(), (C.c_int * 3)()
            capi.check(l.opf_expr_prepare(c["sig"].encode(), F, nf, which, C.byref(r), loc))
            assert rng(r, 2) == c[key] and list(loc)[:2] == c["eloc"], (c["sig"], key)


@pytest.mark.parametrize("mode,exact", [(capi.MODE_EXACT, True), (capi.MODE_FAST, False)])
def test_explicit_fixtures(engine, mode, exact):
    """FTCS2D / FTCS3D / CONV1D runs of the reference itself (examples/FTCS2D/FTCS-OMP.cpp:26, examples/CONV1D/CONV1D.cpp:29-31)"""
    host.set_mode(mode)
    man = json.load(open(os.path.join(GOLD, "manifest.json")))
    for name, info in man.items():
        if "_mpi" in name or "_fbc" in name:  # decomposed set-up / functor BCs: driven through the C++ front-end (tests/test_gpu_frontend.py, tests/test_gpu_multi.py)
            continue
        a = dict(zip(info["args"][::2], info["args"][1::2]))
        case, n, steps, init = a["--case"], int(a["--n"]), int(a["--steps"]), a.get("--init", "zero")
        s, e, ref = O.read_opfd(os.path.join(GOLD, name + ".opfd"))
        if case.startswith("ftcs"):
            dim = 2 if case == "ftcs2d" else 3
            u = build([n] * dim, [0.0] * dim, [1.0] * dim, None, 0, {(d, k): (capi.BC_DIRC, 1.0) for d in range(dim) for k in range(2)}, 0)
            if init == "sin":
                u.initBy(lambda x: float(np.prod([np.sin(np.pi * xi) for xi in x])))
            else:
                u.assign(0.0)
            c = 0.1 / (n - 1) ** 2 * 1.0
            lap = d2x(D2, u) + d2y(D2, u) if dim == 2 else d2x(D2, u) + d2y(D2, u) + d2z(D2, u)
            expr = u + c * lap
        else:
            u = build([n], [0.0], [1.0], None, 0, {(0, 0): (capi.BC_DIRC, 0.0), (0, 1): (capi.BC_DIRC, 0.0)}, 3)
            u.initBy((lambda x: np.sin(2 * np.pi * x[0])) if init == "sin" else (lambda x: 1.0 if 0.2 <= x[0] <= 0.4 else 0.0))
            c = 0.5 / (n - 1) * 1.0
            expr = {"weno_down": u - c * dx(D1WENO53Downwind, u), "weno_up": u + c * dx(D1WENO53Upwind, u),
                    "upwind1": u - c * dx(D1FirstOrderBiasedDownwind, u)}[case]
        for _ in range(steps):
            u.assign(expr)
        got = u.to_numpy(capi.Range.make(s, e))
        if exact:
            assert np.array_equal(got, ref), f"{name}: max abs diff {np.abs(got - ref).max()}"
        else:
            assert np.abs(got - ref).max() <= (1e-11 if "weno" in case else 1e-12) * max(np.abs(ref).max(), 1e-300), name
