"""Coefficient-level parity of the matrix-free operator: probing A = lhs - rhs with unit vectors must reproduce, bit for bit, the
CSR matrices the reference's own test asserts (test/Core/Equation/CSRMatrixGeneratorTest.cpp:57-158: equation
`1.0 == d2x(e) + d2y(e)` on 4x4 cells, Dirichlet / Neumann / periodic; x-fastest row order).  The reference's CSR route pins the
LAST row when asked to (CSRMatrixGenerator.hpp:78,84-91): that row is excluded here (the struct route of the engine pins the
first cell, HYPREEqnSolveHandler.hpp:145-163, checked in test_gpu_implicit.py)."""
import numpy as np
import pytest

from opflow_b200 import capi, host
from opflow_b200.host import D2SecondOrderCentered as D2, d2x, d2y

pytestmark = pytest.mark.gpu

from csr_golden import GOLD, dense


@pytest.mark.parametrize("bc", ["Dirc", "Neum", "Periodic"])
@pytest.mark.parametrize("mode", [capi.MODE_EXACT, capi.MODE_FAST])
def test_probed_operator_equals_reference_csr(engine, bc, mode):
    host.set_mode(mode)
    mesh = host.MeshBuilder(2).newMesh(5, 5).setMeshOfDim(0, 0., 4.).setMeshOfDim(1, 0., 4.).build()

    def mk(name):
        b = host.ExprBuilder().setMesh(mesh).setName(name).setLoc([1, 1]).setExt(1)
        for d in range(2):
            for s in range(2):
                if bc == "Periodic":
                    b.setBC(d, s, host.BCType.Periodic)
                else:
                    b.setBC(d, s, host.BCType.Dirc if bc == "Dirc" else host.BCType.Neum, 0.)
        return b.build()

    e, q = mk("e"), mk("q")
    ar = e.assignableRange
    assert ar.tup(2) == ((0, 0), (4, 4))
    lap = d2x(D2, e) + d2y(D2, e)
    cols = []
    for j in range(16):  # x-fastest numbering of the reference's mapper
        unit = np.zeros((4, 4), order="F")
        unit[j % 4, j // 4] = 1.0
        e.from_numpy(unit)  # upload + updatePadding: ghosts folded by the (homogeneous) BCs, like StencilField does symbolically
        q.assign(lap)
        cols.append(q.to_numpy(ar).reshape(-1, order="F"))
    A = -np.stack(cols, axis=1)  # residual form of `1 == L(e)`: 1 - L(e), the reference stores the negated operator
    G = dense(GOLD[bc])
    rows = slice(0, 15) if GOLD[bc]["pinned_last"] else slice(0, 16)
    assert np.array_equal(A[rows], G[rows]), (A[rows] - G[rows])
    # sparsity pattern (column indices per row) as the reference lists it
    for r in range(16)[rows]:
        assert sorted(np.nonzero(A[r])[0].tolist()) == sorted(GOLD[bc]["col"][GOLD[bc]["ptr"][r]:GOLD[bc]["ptr"][r + 1]])
    # right-hand side: b = rhs - lhs(0) with lhs = 1 (constant) - L(e): -1 on every unpinned row
    e.assign(0.0)
    q.assign(lap)
    rhs = -(1.0 - q.to_numpy(ar).reshape(-1, order="F"))
    assert np.array_equal(rhs[rows], np.asarray(GOLD[bc]["rhs"], dtype=float)[rows])


def _stretched(n, scale):
    s = np.arange(n) / (n - 1)
    return scale * (s + 0.15 * np.sin(2 * np.pi * s) / (2 * np.pi))


def _ref_csr_cases():
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_csr.json")
    return json.load(open(path))["cases"]


@pytest.mark.parametrize("case", _ref_csr_cases(), ids=lambda c: f"{'stretched' if c['stretched'] else 'uniform'}-{'center' if c['loc'] else 'corner'}-bc{c['bc']}")
def test_probed_operator_on_non_unit_spacing(engine, case):
    """Same probe on meshes whose spacing is NOT 1 (7 x 6 nodes on [0,0.7] x [0,1.3], uniform and stretched) against CSR matrices
    assembled by the unmodified reference (oracle/ref_drivers/ref_csr.cpp -> tests/golden/ref_csr.json, hex doubles).
    The reference's implicit path divides a StencilPad by multiplying with the reciprocal (StencilPad.hpp:293-296), its explicit path
    divides: the two differ by one rounding inside the reference itself.  So the statement tested is:
      STENCIL mode  coefficients and right-hand side bit-identical to the reference's assembled matrix;
      EXACT mode    bit-identical to the reference's EXPLICIT operator (test_gpu_explicit.py), hence within 4 ulp of the matrix;
      FAST mode     within 1e-12 relative (reciprocal coefficient arrays, FMA)."""
    c = case
    nx, ny = c["dims"]
    mb = host.MeshBuilder(2).newMesh(nx, ny)
    if c["stretched"]:
        mb.setMeshOfDim(0, _stretched(nx, 0.7)).setMeshOfDim(1, _stretched(ny, 1.3))
    else:
        mb.setMeshOfDim(0, 0., 0.7).setMeshOfDim(1, 0., 1.3)
    mesh = mb.build()

    def mk(name, homogeneous):
        b = host.ExprBuilder().setMesh(mesh).setName(name).setLoc([c["loc"]] * 2).setExt(1)
        for d in range(2):
            if c["bc"] == 0:
                b.setBC(d, 0, host.BCType.Dirc, 0. if homogeneous else 0.25 * (d + 1)).setBC(d, 1, host.BCType.Dirc, 0. if homogeneous else -0.5)
            elif c["bc"] == 1:
                b.setBC(d, 0, host.BCType.Neum, 0. if homogeneous else 0.125).setBC(d, 1, host.BCType.Neum, 0.)
            else:
                b.setBC(d, 0, host.BCType.Periodic).setBC(d, 1, host.BCType.Periodic)
        return b.build()

    (s0, s1), (e0, e1) = c["range"]
    n0, n1 = e0 - s0, e1 - s1
    N = n0 * n1
    G = np.zeros((N, N))
    val = [float.fromhex(v) for v in c["val"]]
    for r in range(N):
        for k in range(c["ptr"][r], c["ptr"][r + 1]):
            G[r, c["col"][k]] = val[k]
    grhs = np.array([float.fromhex(v) for v in c["rhs"]])
    rows = slice(0, N - 1) if c["pinned_last"] else slice(0, N)
    ar = capi.Range.make([s0, s1], [e0, e1])
    for mode in (capi.MODE_STENCIL, capi.MODE_EXACT, capi.MODE_FAST):
        host.set_mode(mode)
        e, q, er = mk("e", True), mk("q", True), mk("er", False)
        assert e.assignableRange.tup(2) == ((s0, s1), (e0, e1))
        lap = d2x(D2, e) + d2y(D2, e)
        cols = []
        for j in range(N):
            unit = np.zeros((n0, n1), order="F")
            unit[j % n0, j // n0] = 1.0
            e.assign(0.0)
            e.from_numpy(unit, ar)
            q.assign(lap)
            cols.append(q.to_numpy(ar).reshape(-1, order="F"))
        A = -np.stack(cols, axis=1)
        er.assign(0.0)  # e = 0 with the REAL boundary data: b = rhs - lhs(0) (generateb, HYPREEqnSolveHandler.hpp:125-179)
        q.assign(d2x(D2, er) + d2y(D2, er))
        rhs = -(1.0 - q.to_numpy(ar).reshape(-1, order="F"))
        scale = np.abs(G).max()
        if mode == capi.MODE_STENCIL:
            assert np.array_equal(A[rows], G[rows]), f"STENCIL mode: {np.count_nonzero(A[rows] != G[rows])} coefficients differ from the reference CSR"
            assert np.array_equal(rhs[rows], grhs[rows]), "STENCIL mode: right-hand side differs from the reference"
            for r in range(N)[rows]:  # sparsity pattern, explicit zeros of the reference included only if it stores them
                assert set(np.nonzero(A[r])[0].tolist()) <= set(c["col"][c["ptr"][r]:c["ptr"][r + 1]])
        elif mode == capi.MODE_EXACT:
            assert np.abs(A[rows] - G[rows]).max() <= 4 * np.finfo(float).eps * scale
            assert np.abs(rhs[rows] - grhs[rows]).max() <= 16 * np.finfo(float).eps * max(1.0, np.abs(grhs).max())
        else:
            assert np.abs(A[rows] - G[rows]).max() <= 1e-12 * scale
            assert np.abs(rhs[rows] - grhs[rows]).max() <= 1e-12 * max(1.0, np.abs(grhs).max())
    host.set_mode(capi.MODE_FAST)


def _csr_dense(ptr, col, val, n):
    a = np.zeros((n, n))
    for r in range(n):
        for k in range(ptr[r], ptr[r + 1]):
            a[r, col[k]] = val[k]
    return a


@pytest.mark.parametrize("bc", ["Dirc", "Neum", "Periodic"])
def test_export_csr_equals_reference_generator(engine, bc):
    """opf_solver_export_csr against the arrays the reference's own test asserts (CSRMatrixGeneratorTest.cpp:57-158): row pointers,
    ascending column indices, values and right-hand side, last row pinned where the reference pins it.  The equation is written
    exactly like the reference's, `1.0 == d2x(e) + d2y(e)` -> residual form lhs = 1 - L(e), rhs = 0."""
    host.set_mode(capi.MODE_EXACT)
    mesh = host.MeshBuilder(2).newMesh(5, 5).setMeshOfDim(0, 0., 4.).setMeshOfDim(1, 0., 4.).build()
    b = host.ExprBuilder().setMesh(mesh).setName("p").setLoc([1, 1]).setExt(1)
    for d in range(2):
        for s in range(2):
            if bc == "Periodic":
                b.setBC(d, s, host.BCType.Periodic)
            else:
                b.setBC(d, s, host.BCType.Dirc if bc == "Dirc" else host.BCType.Neum, 0.)
    p = b.build()
    g = GOLD[bc]
    h = host.EqnSolveHandler(lambda e: (1.0 - (d2x(D2, e) + d2y(D2, e)), 0.0), p, tol=1e-10, maxIter=10)
    ptr, col, val, rhs = h.export_csr(pin_last=g["pinned_last"])
    assert ptr.tolist() == g["ptr"]
    assert col.tolist() == g["col"]
    assert val.tolist() == [float(v) for v in g["val"]]
    assert rhs.tolist() == [float(v) for v in g["rhs"]]
    host.set_mode(capi.MODE_FAST)


@pytest.mark.parametrize("case", _ref_csr_cases(), ids=lambda c: f"{'stretched' if c['stretched'] else 'uniform'}-{'center' if c['loc'] else 'corner'}-bc{c['bc']}")
def test_export_csr_non_unit_spacing_bit_exact(engine, case):
    """Exported CSR in STENCIL mode == the matrix assembled by the unmodified reference on non-unit / stretched spacing
    (tests/golden/ref_csr.json): every stored value and the right-hand side bit for bit; the reference may additionally store explicit
    zeros (a coefficient that cancelled), which a probe cannot see, so the pattern is compared as exported <= reference."""
    c = case
    nx, ny = c["dims"]
    mb = host.MeshBuilder(2).newMesh(nx, ny)
    if c["stretched"]:
        mb.setMeshOfDim(0, _stretched(nx, 0.7)).setMeshOfDim(1, _stretched(ny, 1.3))
    else:
        mb.setMeshOfDim(0, 0., 0.7).setMeshOfDim(1, 0., 1.3)
    mesh = mb.build()
    b = host.ExprBuilder().setMesh(mesh).setName("p").setLoc([c["loc"]] * 2).setExt(1)
    for d in range(2):
        if c["bc"] == 0:
            b.setBC(d, 0, host.BCType.Dirc, 0.25 * (d + 1)).setBC(d, 1, host.BCType.Dirc, -0.5)
        elif c["bc"] == 1:
            b.setBC(d, 0, host.BCType.Neum, 0.125).setBC(d, 1, host.BCType.Neum, 0.)
        else:
            b.setBC(d, 0, host.BCType.Periodic).setBC(d, 1, host.BCType.Periodic)
    host.set_mode(capi.MODE_STENCIL)
    p = b.build()
    (s0, s1), (e0, e1) = c["range"]
    N = (e0 - s0) * (e1 - s1)
    # The reference assembles `1 == L(e)` as 1 - L(e): matrix -L, right-hand side -(1 - L(0)).  Exported here from  L(e) == 1  (no
    # e-free term on the left, so no constant is subtracted from the probes): the same numbers with the opposite sign -- negation
    # is exact --, except the pinned identity row, which is +1 / 0 on both sides.
    h = host.EqnSolveHandler(lambda e: (d2x(D2, e) + d2y(D2, e), 1.0), p, tol=1e-10, maxIter=10)
    ptr, col, val, rhs = h.export_csr(pin_last=c["pinned_last"])
    val, rhs = -val, -rhs
    if c["pinned_last"]:
        val[ptr[N - 1]:ptr[N]] *= -1.0
        rhs[N - 1] *= -1.0
    gval = [float.fromhex(v) for v in c["val"]]
    grhs = np.array([float.fromhex(v) for v in c["rhs"]])
    assert len(ptr) == N + 1
    for r in range(N):
        gold = {c["col"][k]: gval[k] for k in range(c["ptr"][r], c["ptr"][r + 1])}
        mine = {int(col[k]): float(val[k]) for k in range(ptr[r], ptr[r + 1])}
        assert list(mine) == sorted(mine), "columns not ascending"
        assert set(mine) <= set(gold), (r, mine, gold)
        for j, v in gold.items():
            assert mine.get(j, 0.0) == v, (r, j, mine.get(j), v)
    assert np.array_equal(rhs, grhs)
    host.set_mode(capi.MODE_FAST)
