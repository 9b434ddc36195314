#!/usr/bin/env python
"""tools/stress_c4.py [repeats] [n] -- fresh fields + fresh solver + one C4-shaped Poisson solve, repeated; prints every solve that does
not converge in 30 iterations (hunting an intermittent non-convergence seen twice in the GPU suite)."""
import gc
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opflow_b200 import capi, host  # noqa: E402
from opflow_b200.host import D2SecondOrderCentered as D2, EqnSolveHandler, StructSolverType as ST, d2x, d2y  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4097
capi.check(capi.lib().opf_init(0))
host.set_mode(capi.MODE_FAST)
lap = lambda f: d2x(D2, f) + d2y(D2, f)  # noqa: E731
bad = 0
rng = np.random.default_rng(0)
for it in range(reps):
    # churn: a few small fields of random shapes come and go between the solves, like a test-suite does
    junk = []
    for _ in range(int(rng.integers(0, 6))):
        dims = [int(rng.integers(2, 40)) for _ in range(int(rng.integers(2, 4)))]
        mb = host.MeshBuilder(len(dims)).newMesh(*dims)
        for d in range(len(dims)):
            mb.setMeshOfDim(d, 0., 1.)
        b = host.ExprBuilder().setMesh(mb.build()).setName("j").setLoc([1] * len(dims)).setExt(1)
        for d in range(len(dims)):
            b.setBC(d, 0, host.BCType.Neum, 0.3).setBC(d, 1, host.BCType.Dirc, 0.1)
        junk.append(b.build())
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
    bf.assign(lap(pt))
    p.assign(0.0)
    h = EqnSolveHandler(lambda e: (lap(e), bf), p, type_=ST.PCG, precond=ST.PFMG, tol=1e-10, maxIter=100, pinValue=True, staticMat=True)
    st = h.solve()
    if not (st.relerr <= 1e-10 and st.niter <= 30):
        bad += 1
        bsum = host.rangeReduce(bf, capi.RED_SUM)
        bmax = host.rangeReduce(bf, capi.RED_ABSMAX)
        st2 = h.solve()
        print(f"BAD solve {it}: niter {st.niter} relerr {st.relerr:.3e} levels {h.levels()} sum(b) {bsum:.3e} max|b| {bmax:.3e}; solving again: niter {st2.niter} relerr {st2.relerr:.3e}", flush=True)
    del h, p, bf, pt, junk
    if it % 3 == 0:
        gc.collect()
print(f"STRESS {reps} solves at n={n}: {bad} bad", flush=True)
