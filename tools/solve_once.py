#!/usr/bin/env python
"""tools/solve_once.py -- config C4's Poisson handler: set-up, two warm-up solves, then ONE solve between two marker launches;
run under `ncu --metrics gpu__time_duration.sum` to list the launches of a solve (profiles/*_c4_launches.csv)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opflow_b200 import capi, host  # noqa: E402
from opflow_b200.host import D2SecondOrderCentered as D2, EqnSolveHandler, StructSolverType as ST, d2x, d2y  # noqa: E402

l = capi.lib()
capi.check(l.opf_init(0))
host.set_mode(capi.MODE_FAST)
n = int(os.environ.get("OPF_SOLVE_N", "4097"))
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
h = EqnSolveHandler(lambda e: (lap(e), bf), p, type_=ST.PCG, precond=ST.PFMG, tol=1e-10, maxIter=100, pinValue=True, staticMat=True)
for _ in range(2):
    p.assign(0.0)
    st = h.solve()
capi.check(l.opf_synchronize())
print("launches before the profiled solve:", l.opf_launch_count(), flush=True)
p.assign(0.0)
st = h.solve()
capi.check(l.opf_synchronize())
print("iterations", st.niter, "relres", st.relerr, "launches", l.opf_launch_count())
