import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo')
from opflow_b200 import capi, host
from opflow_b200.host import *
l = capi.lib(); capi.check(l.opf_init(0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 513
steps = 20
mesh = host.MeshBuilder(3).newMesh(n, n, n).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).setMeshOfDim(2, 0., 1.).build()
def mk(name, bc=True):
    b = host.ExprBuilder().setName(name).setMesh(mesh)
    if bc:
        for d in range(3):
            b.setBC(d, 0, host.BCType.Dirc, 1.).setBC(d, 1, host.BCType.Dirc, 1.)
    return b.build()
u = mk("u"); v = mk("v"); w = mk("w")
u.assign(1.0); v.assign(2.0); w.assign(0.0)
D2 = D2SecondOrderCentered
cases = {
 "copy w=u (16B)": (w, u + 0 if False else None, "F<0>", [u], [], 16),
 "add w=u+v (24B)": (w, None, "Add<F<0>,F<1>>", [u, v], [], 24),
 "axpy w=u+c*v (24B)": (w, None, "Add<F<0>,Mul<S<0>,F<1>>>", [u, v], [0.5], 24),
 "ftcs3d w=u+c*lap(u) (16B, no alias)": (w, None, "Add<F<0>,Mul<S<0>,Add<Add<D2C<0,F<1>>,D2C<1,F<2>>>,D2C<2,F<3>>>>>", [u,u,u,u], [1e-7], 16),
 "ftcs3d u=u+c*lap(u) (16B, alias)": (u, None, "Add<F<0>,Mul<S<0>,Add<Add<D2C<0,F<1>>,D2C<1,F<2>>>,D2C<2,F<3>>>>>", [u,u,u,u], [1e-7], 16),
 "lap w=lap(u) (16B)": (w, None, "Add<Add<D2C<0,F<0>>,D2C<1,F<1>>>,D2C<2,F<2>>>", [u,u,u], [], 16),
 "d2z w=d2z(u) (16B)": (w, None, "D2C<2,F<0>>", [u], [], 16),
 "d2x w=d2x(u) (16B)": (w, None, "D2C<0,F<0>>", [u], [], 16),
 "d2y w=d2y(u) (16B)": (w, None, "D2C<1,F<0>>", [u], [], 16),
 "lapxy w=d2x+d2y (16B)": (w, None, "Add<D2C<0,F<0>>,D2C<1,F<1>>>", [u,u], [], 16),
}
ms = C.c_float()
for name, (dst, _, sig, fs, sc, bpc) in cases.items():
    F = (C.c_void_p * len(fs))(*[f.h for f in fs]); S = (C.c_double * max(1,len(sc)))(*sc)
    for _ in range(3): capi.check(l.opf_assign(dst.h, 0, sig.encode(), F, len(fs), S, len(sc)))
    capi.check(l.opf_synchronize()); capi.check(l.opf_timer_begin())
    for _ in range(steps): capi.check(l.opf_assign(dst.h, 0, sig.encode(), F, len(fs), S, len(sc)))
    capi.check(l.opf_timer_end(C.byref(ms)))
    t = ms.value / steps
    cells = (n-2)**3
    print(f"{name:45s} {t:.4f} ms  {cells*bpc/t/1e6:8.1f} GB/s  {cells/t/1e6:7.1f} GLUPS")
