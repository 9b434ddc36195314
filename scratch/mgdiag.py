import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from opflow_b200 import capi, host
from opflow_b200.host import *
from test_gpu_implicit import build, lap, ST
l = capi.lib(); capi.check(l.opf_init(0))
host.set_mode(capi.MODE_FAST)
n = 65
for bc in ("Neum", "Periodic", "Dirc"):
  for pin in (True, False):
    c = {"n": [n, n], "lo": [0, 0], "hi": [1, 1], "loc": [1,1], "bc": bc, "bcv": 0.0, "ext": 1}
    p, bf, pt = build(c, "p"), build(c, "b"), build(c, "pt")
    sh = pt.localRange.shape(2)
    xs = [ (np.arange(sh[d]) + 0.5) / (n - 1) for d in range(2)]
    pt.from_numpy(np.asfortranarray(np.cos(2*np.pi * xs[0])[:, None] * np.cos(2 * np.pi * xs[1])[None, :]))
    bf.assign(lap(pt, 2))
    res = []
    for k in (1, 2, 3, 4, 6, 8):
        p.assign(0.0)
        h = EqnSolveHandler(lambda e: (lap(e, 2), bf), p, type_=ST.PFMG, precond=ST.NONE, tol=1e-14, maxIter=k, pinValue=pin, numPreRelax=2, numPostRelax=2)
        st = h.solve(); res.append(st.relerr)
    print(bc, "pin", pin, " ".join(f"{r:.2e}" for r in res), "bsum", host.rangeReduce(bf, 0))
