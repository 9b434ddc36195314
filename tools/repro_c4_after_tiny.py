#!/usr/bin/env python
"""tools/repro_c4_after_tiny.py [maxIter] [n] -- the ten tiny cell-centred edge-case tests followed, in the same process, by the C4-shaped
Poisson solve (the sequence that made test_c4_poisson_solution_satisfies_the_explicit_operator fail): prints the solve state."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
maxit = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4097
os.chdir(ROOT)
pytest.main(["tests/test_gpu_edge_cases.py", "-q", "-m", "gpu", "-k", "cell_centred"])
from opflow_b200 import capi, host  # noqa: E402
from opflow_b200.host import D2SecondOrderCentered as D2, EqnSolveHandler, StructSolverType as ST, d2x, d2y  # noqa: E402

host.set_mode(capi.MODE_FAST)
mesh = host.MeshBuilder(2).newMesh(n, n).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).build()


def mk(name):
    b = host.ExprBuilder().setMesh(mesh).setName(name).setLoc([1, 1]).setExt(1)
    for d in range(2):
        b.setBC(d, 0, host.BCType.Neum, 0.).setBC(d, 1, host.BCType.Neum, 0.)
    return b.build()


p, bf, pt = mk("p"), mk("b"), mk("pt")
sh = pt.localRange.shape(2)
xs = [(np.arange(sh[d]) + 0.5) / (n - 1) for d in range(2)]
pt.from_numpy(np.asfortranarray(np.cos(2 * np.pi * xs[0])[:, None] * np.cos(np.pi * xs[1])[None, :]))
lap = lambda f: d2x(D2, f) + d2y(D2, f)  # noqa: E731
bf.assign(lap(pt))
p.assign(0.0)
h = EqnSolveHandler(lambda e: (lap(e), bf), p, type_=ST.PCG, precond=ST.PFMG, tol=1e-10, maxIter=maxit, pinValue=True, staticMat=True)
st = h.solve()
print("REPRO niter", st.niter, "relerr", st.relerr, "levels", h.levels(), flush=True)
