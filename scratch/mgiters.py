import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from opflow_b200 import capi, host
from opflow_b200.host import *
from test_gpu_implicit import build, lap, ST
import ctypes as C, time
l = capi.lib(); capi.check(l.opf_init(0))
host.set_mode(capi.MODE_FAST)
for bc, loc, pin in (("Neum",[1,1],True), ("Dirc",[1,1],False), ("Dirc",[0,0],False), ("Periodic",[1,1],True)):
  for n in (65, 257, 1025, 4097):
    c = {"n": [n, n], "lo": [0, 0], "hi": [1, 1], "loc": loc, "bc": bc, "bcv": 0.0, "ext": 1}
    p, bf, pt = build(c, "p"), build(c, "b"), build(c, "pt")
    r = pt.localRange; sh = r.shape(2)
    xs = [ (np.arange(sh[d]) + (0.5 if loc[d] else 0.0)) / (n - 1) for d in range(2)]
    pt.from_numpy(np.asfortranarray(np.cos(2*np.pi * xs[0])[:, None] * np.sin(2 * np.pi * xs[1])[None, :]))
    bf.assign(lap(pt, 2)); p.assign(0.0)
    for (ty, pc) in ((ST.PCG, ST.PFMG), (ST.PFMG, ST.NONE)):
        p.assign(0.0)
        h = EqnSolveHandler(lambda e: (lap(e, 2), bf), p, type_=ty, precond=pc, tol=1e-10, maxIter=100, pinValue=pin, staticMat=True, numPreRelax=2, numPostRelax=2)
        st = h.solve()   # includes setup
        p.assign(0.0); host.synchronize(); t0 = time.perf_counter(); st = h.solve(); host.synchronize(); t1 = time.perf_counter()
        print(f"{bc:8s} loc={loc} n={n:5d} solver={ty} pc={pc} levels={h.levels()} iters={st.niter} relres={st.relerr:.2e} solve_ms={(t1-t0)*1e3:.2f}")
