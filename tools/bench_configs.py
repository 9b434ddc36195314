#!/usr/bin/env python
"""tools/bench_configs.py -- the other BASELINE.json configurations, one JSON line each (bench.py itself reports C2 only).

C1  FTCS2D 1025^2 (examples/FTCS2D/FTCS-OMP.cpp)      window_kernel (2-D register-window skeleton), 16 B / update
C2x FTCS3D 513^3 in EXACT mode (bit-identical arithmetic: true divides, no FMA)   tma_kernel, FP64-divide-bound
C3  CONV1D WENO5 2^26 cells                           assign_kernel (1-D), 16 B / update, FP64-pipe-bound
C4  pressure Poisson 4097^2 (LidDriven2D.cpp:67-74)   PCG + multigrid, ms per solve
All timed with CUDA events on the engine stream after warm-up; fields are larger than L2 except C1 (8 MB: L2-resident, noted).
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opflow_b200 import capi, host  # noqa: E402
from opflow_b200.host import (D1WENO53Downwind, D2SecondOrderCentered as D2, EqnSolveHandler, StructSolverType as ST, d2x, d2y, d2z, dx)  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
l = capi.lib()
capi.check(l.opf_init(0))


def timed(fn, steps, warmup=5, batches=3):
    """best of `batches` timed batches: the first batch after a previous workload's fields were freed in the same process runs up to
    20x slower for a few dozen launches (fresh allocations; seen on the 1-D case only after 2-D/3-D cases) -- a transient of the
    measuring script, not of the kernels"""
    for _ in range(warmup):
        fn()
    best = None
    for _ in range(batches):
        capi.check(l.opf_synchronize())
        ms = C.c_float()
        capi.check(l.opf_timer_begin())
        for _ in range(steps):
            fn()
        capi.check(l.opf_timer_end(C.byref(ms)))
        t = ms.value / steps
        best = t if best is None else min(best, t)
    return best


def explicit(name, u, expr, updates, steps, note=""):
    ms = timed(lambda: u.assign(expr), steps)
    gbs = 16.0 * updates / (ms * 1e-3) / 1e9
    print(json.dumps({"config": name, "ms_per_step": ms, "glups": updates / (ms * 1e-3) / 1e9, "achieved_gbs": gbs, "hbm_frac": gbs / PEAK,
                      "updates_per_step": updates, "note": note}), flush=True)


def dirichlet(dim, dims, bcv, ext=0):
    mb = host.MeshBuilder(dim).newMesh(*dims)
    for d in range(dim):
        mb.setMeshOfDim(d, 0., 1.)
    b = host.ExprBuilder().setName("u").setMesh(mb.build()).setExt(ext)
    for d in range(dim):
        b.setBC(d, 0, host.BCType.Dirc, bcv).setBC(d, 1, host.BCType.Dirc, bcv)
    return b.build()


host.set_mode(capi.MODE_FAST)
n = 1025
u = dirichlet(2, (n, n), 1.0)
u.assign(0.0)
explicit("C1 FTCS2D 1025^2 fast", u, u + (0.1 / (n - 1) ** 2) * (d2x(D2, u) + d2y(D2, u)), (n - 2) ** 2, 2000, "8.4 MB field: L2-resident, launch-latency bound")
n = 4097
u = dirichlet(2, (n, n), 1.0)
u.assign(0.0)
explicit("C1' FTCS2D 4097^2 fast", u, u + (0.1 / (n - 1) ** 2) * (d2x(D2, u) + d2y(D2, u)), (n - 2) ** 2, 500, "134 MB field (> L2)")
del u
n = 513
host.set_mode(capi.MODE_EXACT)
u = dirichlet(3, (n, n, n), 1.0)
u.assign(0.0)
explicit("C2 FTCS3D 513^3 EXACT", u, u + (0.1 / (n - 1) ** 2) * (d2x(D2, u) + d2y(D2, u) + d2z(D2, u)), (n - 2) ** 3, 50, "9 IEEE divides per cell: FP64-pipe bound")
del u
n = 2 ** 26 + 1
for mode, nm in ((capi.MODE_FAST, "fast"), (capi.MODE_EXACT, "EXACT")):
    host.set_mode(mode)
    u = dirichlet(1, (n,), 0.0, ext=3)
    x = np.linspace(0., 1., n)
    u.from_numpy(np.where((x >= 0.2) & (x <= 0.4), 1.0, 0.0))
    explicit(f"C3 WENO5 2^26 {nm}", u, u - (0.5 / (n - 1)) * dx(D1WENO53Downwind, u), n - 2, 50, "FP64-pipe bound (~20 divides per cell in EXACT)")
    del u
host.set_mode(capi.MODE_FAST)
for n in (1025, 4097):
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
    lap = lambda f: d2x(D2, f) + d2y(D2, f)
    bf.assign(lap(pt))
    h = EqnSolveHandler(lambda e: (lap(e), bf), p, type_=ST.PCG, precond=ST.PFMG, tol=1e-10, maxIter=100, pinValue=True, staticMat=True)
    state = {}

    def solve():
        p.assign(0.0)
        state["st"] = h.solve()

    ms = timed(solve, 5, 2, 2)
    st = state["st"]
    cells = (n - 1) ** 2
    print(json.dumps({"config": f"C4 Poisson {n - 1}^2 cells, Neumann + pin, PCG + GMG V(1,1), tol 1e-10", "ms_per_solve": ms, "iterations": st.niter,
                      "relres": st.relerr, "levels": h.levels(), "cells_x_iterations_per_s": cells * st.niter / (ms * 1e-3),
                      "effective_gbs_at_200B_per_cell_iter": 200.0 * cells * st.niter / (ms * 1e-3) / 1e9}), flush=True)
    del h, p, bf, pt
