"""Implicit path: matrix-free PCG / BiCGSTAB / multigrid on the device vs solves by the UNMODIFIED reference
(HYPREEqnSolveHandler + HYPRE GMRES/PFMG, tests/golden/ref_implicit.json from oracle/ref_drivers/ref_implicit.cpp).
Both sides are driven to ~1e-13 relative residual; the solutions must agree to 1e-10 relative (north_star tolerance)."""
import json
import os

import numpy as np
import pytest

from opflow_b200 import capi, host
from opflow_b200.host import D2SecondOrderCentered as D2, EqnSolveHandler, StructSolverType as ST, d2x, d2y, d2z

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = json.load(open(os.path.join(GOLD, "ref_implicit.json")))["cases"]
BCT = {"Dirc": capi.BC_DIRC, "Neum": capi.BC_NEUM, "Periodic": capi.BC_PERIODIC}


def build(c, name, bcv=None):
    dim = len(c["n"])
    mb = host.MeshBuilder(dim).newMesh(*c["n"])
    for d in range(dim):
        mb.setMeshOfDim(d, c["lo"][d], c["hi"][d])
    b = host.ExprBuilder().setMesh(mb.build()).setName(name).setLoc(c["loc"]).setExt(c["ext"])
    for d in range(dim):
        for s in range(2):
            b.setBC(d, s, BCT[c["bc"]], c["bcv"] if bcv is None else bcv)
    return b.build()


def lap(e, dim):
    return d2x(D2, e) + d2y(D2, e) if dim == 2 else d2x(D2, e) + d2y(D2, e) + d2z(D2, e)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
@pytest.mark.parametrize("solver,precond", [(ST.PCG, ST.PFMG), (ST.PCG, ST.Jacobi), (ST.BICGSTAB, ST.PFMG), (ST.GMRES, ST.NONE), (ST.PFMG, ST.NONE)])
def test_poisson_matches_reference(engine, case, solver, precond):
    host.set_mode(capi.MODE_EXACT)
    dim = len(case["n"])
    p = build(case, "p")
    bf = build(case, "b")
    rng = capi.Range.make(case["range"][0], case["range"][1])
    shape = rng.shape(dim)
    bf.from_numpy(np.array(case["b"]).reshape(shape, order="F"), rng)
    p.assign(0.0)
    h = EqnSolveHandler(lambda e: (lap(e, dim), bf), p, type_=solver, precond=precond, tol=1e-13, maxIter=4000 if precond != ST.PFMG and solver != ST.PFMG else 200,
                        pinValue=bool(case["pin"]), numPreRelax=2, numPostRelax=2)
    if precond == ST.PFMG or solver == ST.PFMG:
        assert h.levels() >= 3, "multigrid hierarchy was not built"
    st = h.solve()
    ref = np.array(case["p"]).reshape(shape, order="F")
    got = p.to_numpy(rng)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert st.relerr <= 1e-12 or st.niter >= 200, (st.niter, st.relerr)
    assert err <= 1e-10, f"{case['name']}: solution differs from the reference by {err:.3e} (niter={st.niter}, relres={st.relerr:.2e})"


def test_multigrid_is_mesh_independent(engine):
    """V-cycle preconditioned CG: iteration count must not grow with resolution (sanity of restriction/prolongation)"""
    host.set_mode(capi.MODE_FAST)
    iters = []
    for n in (65, 129, 257, 513):
        c = {"n": [n, n], "lo": [0, 0], "hi": [1, 1], "loc": [1, 1], "bc": "Neum", "bcv": 0.0, "ext": 1}
        p, bf, pt = build(c, "p"), build(c, "b"), build(c, "pt")
        pt.initBy(lambda x: np.cos(np.pi * x[0]) * np.cos(2 * np.pi * x[1])) if n <= 129 else None
        if n > 129:
            xs = (np.arange(n - 1) + 0.5) / (n - 1)
            pt.from_numpy(np.asfortranarray(np.cos(np.pi * xs)[:, None] * np.cos(2 * np.pi * xs)[None, :]))
        bf.assign(lap(pt, 2))
        p.assign(0.0)
        h = EqnSolveHandler(lambda e: (lap(e, 2), bf), p, type_=ST.PCG, precond=ST.PFMG, tol=1e-10, maxIter=100, pinValue=True, staticMat=True,
                            numPreRelax=2, numPostRelax=2)
        st = h.solve()
        assert st.relerr <= 1e-10
        iters.append(st.niter)
        # solving again with the converged field as the initial guess costs (almost) nothing
        st2 = h.solve()
        assert st2.niter <= 1
    assert max(iters) <= 2 * min(iters) + 4, iters


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
@pytest.mark.parametrize("relax", [2, 3])
def test_red_black_gauss_seidel_smoother(engine, case, relax):
    """PFMG relaxType 2 / 3 (StructSolverPFMG.hpp:23-34: symmetric and plain red-black Gauss-Seidel) as the V-cycle smoother: the
    solution is the reference's (HYPRE solve, 1e-10 relative) whatever the smoother; the symmetric variant keeps PCG applicable,
    the plain one runs under GMRES.  The stronger smoother must not need more iterations than weighted Jacobi."""
    host.set_mode(capi.MODE_EXACT)
    dim = len(case["n"])
    rng = capi.Range.make(case["range"][0], case["range"][1])
    shape = rng.shape(dim)
    ref = np.array(case["p"]).reshape(shape, order="F")
    iters = {}
    for rt in (1, relax):
        p, bf = build(case, "p"), build(case, "b")
        bf.from_numpy(np.array(case["b"]).reshape(shape, order="F"), rng)
        p.assign(0.0)
        h = EqnSolveHandler(lambda e: (lap(e, dim), bf), p, type_=ST.PCG if rt != 3 else ST.GMRES, precond=ST.PFMG, tol=1e-13, maxIter=200,
                            pinValue=bool(case["pin"]), relaxType=rt)
        st = h.solve()
        iters[rt] = st.niter
        err = np.abs(p.to_numpy(rng) - ref).max() / np.abs(ref).max()
        assert err <= 1e-10, f"{case['name']} relaxType {rt}: solution differs from the reference by {err:.3e} (niter={st.niter}, relres={st.relerr:.2e})"
    if relax == 2:
        assert iters[2] <= iters[1], iters


def test_red_black_smoother_at_c4_shape(engine):
    """4-level-deep hierarchy, all-Neumann pinned Poisson problem of LidDriven2D (BASELINE C4's operator) at 1024^2 cells in FAST mode:
    iteration counts of V(1,1) cycles with weighted Jacobi and with symmetric red-black Gauss-Seidel, same solution"""
    host.set_mode(capi.MODE_FAST)
    n = 1025
    c = {"n": [n, n], "lo": [0, 0], "hi": [1, 1], "loc": [1, 1], "bc": "Neum", "bcv": 0.0, "ext": 1}
    xs = (np.arange(n - 1) + 0.5) / (n - 1)
    exact = np.asfortranarray(np.cos(np.pi * xs)[:, None] * np.cos(2 * np.pi * xs)[None, :])
    sols, iters = {}, {}
    for rt in (1, 2):
        p, bf, pt = build(c, "p"), build(c, "b"), build(c, "pt")
        pt.from_numpy(exact)
        bf.assign(lap(pt, 2))
        p.assign(0.0)
        h = EqnSolveHandler(lambda e: (lap(e, 2), bf), p, type_=ST.PCG, precond=ST.PFMG, tol=1e-10, maxIter=100, pinValue=True, staticMat=True, relaxType=rt)
        st = h.solve()
        assert st.relerr <= 1e-10
        sols[rt], iters[rt] = p.to_numpy(), st.niter
    d = sols[1] - sols[2]
    assert np.abs(d - d.mean()).max() <= 1e-8 * np.abs(sols[1]).max()
    assert iters[2] <= iters[1], iters
    print("iterations: Jacobi", iters[1], "red-black GS", iters[2])


@pytest.mark.parametrize("shape,loc", [((70, 37), [1, 1]), ((41, 23), [0, 0]), ((130, 21, 19), [1, 1, 1]), ((71, 17, 13), [0, 1, 0])])
def test_parity_leaf_masks_by_global_index(engine, shape, loc):
    """Par<c> is 1 where (i + j + k) of the GLOBAL cell index has parity c -- also for Corner-Dirichlet fields whose assignable range starts
    at 1, in the register-window skeleton (2-D, small 3-D) and the TMA skeleton (3-D rows >= 64): x += Par<c> * d * r  against numpy"""
    import ctypes as C
    host.set_mode(capi.MODE_EXACT)
    dim = len(shape)
    mb = host.MeshBuilder(dim).newMesh(*shape)
    for d in range(dim):
        mb.setMeshOfDim(d, 0., 1.)
    mesh = mb.build()

    def mk(name):
        b = host.ExprBuilder().setMesh(mesh).setName(name).setLoc(loc).setExt(1)
        for d in range(dim):
            b.setBC(d, 0, host.BCType.Dirc, 0.).setBC(d, 1, host.BCType.Dirc, 0.)
        return b.build()
    x, dv, r = mk("x"), mk("d"), mk("r")
    ar = x.assignableRange
    shp = ar.shape(dim)
    rs = np.random.RandomState(7)
    X, D, R = (np.asfortranarray(rs.rand(*shp)) for _ in range(3))
    idx = np.indices(shp).astype(np.int64)
    gsum = sum(idx[d] + ar.start[d] for d in range(dim))
    for c in (0, 1):
        x.from_numpy(X, ar), dv.from_numpy(D, ar), r.from_numpy(R, ar)
        F = (C.c_void_p * 3)(x.h, dv.h, r.h)
        S = (C.c_double * 1)(0.0)
        capi.check(capi.lib().opf_assign(x.h, 0, f"Add<F<0>,Mul<Par<{c}>,Mul<F<1>,F<2>>>>".encode(), F, 3, S, 0))
        want = X + np.where((gsum & 1) == c, 1.0, 0.0) * (D * R)
        assert np.array_equal(x.to_numpy(ar), want), f"colour {c}"
    host.set_mode(capi.MODE_FAST)
