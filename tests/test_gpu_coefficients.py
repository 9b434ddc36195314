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
