import sys, ctypes as C, numpy as np, os
sys.path.insert(0,'/root/repo')
from opflow_b200 import capi, host
from opflow_b200.host import D2SecondOrderCentered as D2, d2x, d2y
l = capi.lib(); capi.check(l.opf_init(0))
n = 4097
mb = host.MeshBuilder(2).newMesh(n,n).setMeshOfDim(0,0.,1.).setMeshOfDim(1,0.,1.)
b = host.ExprBuilder().setName("u").setMesh(mb.build())
for d in range(2): b.setBC(d,0,host.BCType.Dirc,1.).setBC(d,1,host.BCType.Dirc,1.)
u = b.build(); u.assign(0.0)
e = u + (0.1/(n-1)**2)*(d2x(D2,u)+d2y(D2,u))
for _ in range(10): u.assign(e)
best=1e9
for rep in range(3):
    capi.check(l.opf_synchronize()); ms=C.c_float(); capi.check(l.opf_timer_begin())
    for _ in range(200): u.assign(e)
    capi.check(l.opf_timer_end(C.byref(ms))); best=min(best, ms.value/200)
print(os.environ.get("OPF_WTX"), os.environ.get("OPF_WCH"), os.environ.get("OPF_WPD"), round(best*1e3,2), "us", round(16*(n-2)**2/best/1e6/6453.7,3))
