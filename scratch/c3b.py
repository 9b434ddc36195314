import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo')
from opflow_b200 import capi, host
from opflow_b200.host import D1WENO53Downwind, D2SecondOrderCentered as D2, d2x, d2y, d2z, dx
l = capi.lib(); capi.check(l.opf_init(0))
variant = sys.argv[1]
def timed(fn, steps, warmup=5):
    for _ in range(warmup): fn()
    capi.check(l.opf_synchronize()); ms = C.c_float(); capi.check(l.opf_timer_begin())
    for _ in range(steps): fn()
    capi.check(l.opf_timer_end(C.byref(ms))); return ms.value/steps
def dirichlet(dim, dims, bcv, ext=0):
    mb = host.MeshBuilder(dim).newMesh(*dims)
    for d in range(dim): mb.setMeshOfDim(d, 0., 1.)
    b = host.ExprBuilder().setName("u").setMesh(mb.build()).setExt(ext)
    for d in range(dim): b.setBC(d, 0, host.BCType.Dirc, bcv).setBC(d, 1, host.BCType.Dirc, bcv)
    return b.build()
if "exact3d" in variant:
    host.set_mode(capi.MODE_EXACT)
    n = 257
    u = dirichlet(3, (n,n,n), 1.0); u.assign(0.0)
    print("c2 exact", timed(lambda: u.assign(u + (0.1/(n-1)**2)*(d2x(D2,u)+d2y(D2,u)+d2z(D2,u))), 10), flush=True)
    del u
if "fast2d" in variant:
    host.set_mode(capi.MODE_FAST)
    n = 1025
    u = dirichlet(2, (n,n), 1.0); u.assign(0.0)
    print("c1 fast", timed(lambda: u.assign(u + (0.1/(n-1)**2)*(d2x(D2,u)+d2y(D2,u))), 10), flush=True)
    del u
n = 2**26+1
host.set_mode(capi.MODE_FAST)
u = dirichlet(1, (n,), 0.0, ext=3)
x = np.linspace(0.,1.,n); u.from_numpy(np.where((x>=0.2)&(x<=0.4),1.0,0.0))
e = u - (0.5/(n-1))*dx(D1WENO53Downwind,u)
print(variant, "c3 fast", timed(lambda: u.assign(e), 50), flush=True)
print(variant, "c3 fast again", timed(lambda: u.assign(e), 20), flush=True)
