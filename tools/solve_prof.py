import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from opflow_b200 import capi, host
from opflow_b200.host import *
from test_gpu_implicit import build, lap, ST
import ctypes as C, time
l = capi.lib(); capi.check(l.opf_init(0))
host.set_mode(capi.MODE_FAST)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
bc, loc, pin = "Neum", [1,1], True
c = {"n": [n, n], "lo": [0, 0], "hi": [1, 1], "loc": loc, "bc": bc, "bcv": 0.0, "ext": 1}
p, bf, pt = build(c, "p"), build(c, "b"), build(c, "pt")
r = pt.localRange; sh = r.shape(2)
xs = [ (np.arange(sh[d]) + (0.5 if loc[d] else 0.0)) / (n - 1) for d in range(2)]
pt.from_numpy(np.asfortranarray(np.cos(2*np.pi * xs[0])[:, None] * np.sin(2 * np.pi * xs[1])[None, :]))
bf.assign(lap(pt, 2)); p.assign(0.0)
h = EqnSolveHandler(lambda e: (lap(e, 2), bf), p, type_=ST.PCG, precond=ST.PFMG, tol=1e-10, maxIter=100, pinValue=pin, staticMat=True, numPreRelax=1, numPostRelax=1)
st = h.solve()
for rep in range(3):
    p.assign(0.0); host.synchronize(); l0 = l.opf_launch_count(); t0 = time.perf_counter(); st = h.solve(); host.synchronize(); t1 = time.perf_counter()
    print(f"n={n} levels={h.levels()} iters={st.niter} relres={st.relerr:.2e} solve_ms={(t1-t0)*1e3:.2f} launches={l.opf_launch_count()-l0}")
