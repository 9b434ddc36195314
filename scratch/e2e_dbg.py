import sys, time, ctypes as C, numpy as np, torch
sys.path.insert(0,'/root/repo')
from opflow_b200 import capi, host
from opflow_b200.host import D2SecondOrderCentered as D2, d2x, d2y, d2z
l = capi.lib(); capi.check(l.opf_init(0))
n = 513
mesh = host.MeshBuilder(3).newMesh(n, n, n).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).setMeshOfDim(2, 0., 1.).build()
b = host.ExprBuilder().setName("u").setMesh(mesh)
for d in range(3):
    b.setBC(d, 0, host.BCType.Dirc, 1.).setBC(d, 1, host.BCType.Dirc, 1.)
u = b.build(); u.assign(0.0)
lr = u.localRange
hin = torch.zeros((n, n, n), dtype=torch.float64).pin_memory(); hout = torch.zeros((n, n, n), dtype=torch.float64).pin_memory()
print("pinned:", hin.is_pinned(), hout.is_pinned())
dev = torch.empty((n, n, n), dtype=torch.float64, device="cuda")
def t(f, reps=3):
    f(); torch.cuda.synchronize(); capi.check(l.opf_synchronize())
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); capi.check(l.opf_synchronize())
    return (time.perf_counter() - t0) / reps * 1e3
print("torch dense H2D ms", t(lambda: dev.copy_(hin, non_blocking=True)))
print("torch dense D2H ms", t(lambda: hout.copy_(dev, non_blocking=True)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): dev.copy_(hin, non_blocking=True)
    with torch.cuda.stream(s2): hout.copy_(dev, non_blocking=True)
print("torch dense both ms", t(both))
c = 0.1 / (n - 1) ** 2
expr = u + c * (d2x(D2, u) + d2y(D2, u) + d2z(D2, u))
sig, fields, scalars = expr.flatten()
F = (C.c_void_p * len(fields))(*[f.h for f in fields]); S = (C.c_double * len(scalars))(*scalars)
def step():
    capi.check(l.opf_assign_host(u.h, capi.OP_EQ, sig.encode(), F, len(fields), S, len(scalars), u.h, C.c_void_p(hin.data_ptr()), C.c_void_p(hout.data_ptr())))
print("opf_assign_host ms", t(step, 4))
print("upload pitched ms", t(lambda: u.upload_raw(hin.data_ptr(), lr)))
