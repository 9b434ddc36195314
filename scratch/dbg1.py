import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from opflow_b200 import capi, host
from opflow_b200.host import *
from oracle import oracle as O
from helpers import *
from test_gpu_explicit import stretched
capi.check(capi.lib().opf_init(0))
host.set_mode(0)
dims = [37, 21, 13]
coords = [stretched(n) for n in dims]
bc = {(0, 0): (capi.BC_DIRC, 1.0), (0, 1): (capi.BC_NEUM, 0.5), (1, 0): (capi.BC_NEUM, -0.25), (1, 1): (capi.BC_DIRC, 2.0),
      (2, 0): (capi.BC_SYMM, 0.0), (2, 1): (capi.BC_ASYMM, 0.0)}
g, o = make_pair(dims, None, None, loc=[1, 1, 1], bc=bc, ext=1, coords=coords)
print("ranges", g.localRange.tup(), g.assignableRange.tup(), g.logicalRange.tup(), g.storageRange.tup())
print("oracle", o.localRange, o.assignableRange, o.logicalRange, o.storageRange)
for d in range(3):
    x,dx,idx = g.mesh.axis(d)
    print("mesh eq", d, np.array_equal(x, o.mesh.x[d]), np.array_equal(dx, o.mesh.dx[d]))
set_both(g, o)
def cmp(tag):
    a, b = gpu_storage(g, o)
    diff = np.abs(a-b)
    print(tag, "max", diff.max(), "at", np.unravel_index(diff.argmax(), diff.shape), "nbad", (diff>0).sum(), "shape", a.shape)
    if diff.max()>0:
        bad = np.argwhere(diff>0)
        print(" bad min idx", bad.min(0), "max idx", bad.max(0))
        for ax in range(3):
            print("  axis",ax,"bad planes", sorted(set(bad[:,ax]))[:20])
cmp("after init")
c = 1e-5
D = D2SecondOrderCentered
e = g + c * (d2x(D, g) + d2y(D, g) + d2z(D, g))
g.assign(e); O.assign(o, e.signature(), [o]*4, [c])
cmp("after 1 step")
