import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo')
from opflow_b200 import capi, host
from opflow_b200.host import D1WENO53Downwind, dx
l = capi.lib(); capi.check(l.opf_init(0))
n = 2**26+1
def timed(fn, steps, warmup=5):
    for _ in range(warmup): fn()
    capi.check(l.opf_synchronize()); ms = C.c_float(); capi.check(l.opf_timer_begin())
    for _ in range(steps): fn()
    capi.check(l.opf_timer_end(C.byref(ms))); return ms.value/steps
mesh = host.MeshBuilder(1).newMesh(n).setMeshOfDim(0,0.,1.).build()
for mode,nm in ((capi.MODE_FAST,"fast"),(capi.MODE_EXACT,"exact"),(capi.MODE_FAST,"fast")):
    host.set_mode(mode)
    u = host.ExprBuilder().setMesh(mesh).setName("u").setBC(0,0,host.BCType.Dirc,0.).setBC(0,1,host.BCType.Dirc,0.).setExt(3).build()
    x = np.linspace(0.,1.,n); u.from_numpy(np.where((x>=0.2)&(x<=0.4),1.0,0.0))
    e = u - (0.5/(n-1))*dx(D1WENO53Downwind,u)
    for steps in (5, 20, 50):
        ms = timed(lambda: u.assign(e), steps, 2)
        a = u.to_numpy(); nz = a[a!=0]
        print(nm, steps, round(ms,3), "min|nonzero|", np.abs(nz).min(), "count denormal", int((np.abs(nz) < 2.3e-308).sum()), flush=True)
    del u
