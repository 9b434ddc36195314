#!/usr/bin/env python
"""tools/bench_poisson_mgpu.py -- weak scaling of the pressure Poisson solve (BASELINE C4/C5 implicit part):
    python -m torch.distributed.run --nproc-per-node N tools/bench_poisson_mgpu.py [--size 4097] [--dim 2]
2-D: n x (n-1)*N+1 nodes, cell-centred, Neumann + pinned value (LidDriven2D.cpp:67-74); 3-D: periodic + pin (TaylorGreen box).
y- (z-) slabs, PCG + distributed geometric multigrid, tol 1e-10.  Prints one JSON line on rank 0."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opflow_b200 import capi, host  # noqa: E402
from opflow_b200.host import D2SecondOrderCentered as D2, EqnSolveHandler, StructSolverType as ST, d2x, d2y, d2z  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=4097, help="nodes per axis per GPU")
ap.add_argument("--dim", type=int, default=2)
ap.add_argument("--solves", type=int, default=5)
args = ap.parse_args()
rank, world, lrank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lrank)
l = capi.lib()
capi.check(l.opf_init(lrank))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = (C.c_ubyte * 128)()
        capi.check(l.opf_comm_unique_id(raw))
        idbuf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    dist.broadcast(idbuf, 0)
    capi.check(l.opf_comm_init(rank, world, (C.c_ubyte * 128)(*idbuf.cpu().tolist())))
host.set_mode(capi.MODE_FAST)
n, dim = args.size, args.dim
dims = [n] * (dim - 1) + [(n - 1) * world + 1]
mb = host.MeshBuilder(dim).newMesh(*dims)
for d in range(dim):
    mb.setMeshOfDim(d, 0., 1. if d < dim - 1 else float(world))
mesh = mb.build()


def mk(name):
    b = host.ExprBuilder().setMesh(mesh).setName(name).setLoc([1] * dim).setExt(1)
    for d in range(dim):
        if dim == 3:
            b.setBC(d, 0, host.BCType.Periodic).setBC(d, 1, host.BCType.Periodic)
        else:
            b.setBC(d, 0, host.BCType.Neum, 0.).setBC(d, 1, host.BCType.Neum, 0.)
    if world > 1:
        b.setPadding(1).setSplitStrategy(world, rank, host.split_slab(mesh, world))
    return b.build()


p, bf, pt = mk("p"), mk("b"), mk("pt")
lr = pt.localRange
k = 2 * np.pi
xs = [(np.arange(lr.start[d], lr.end[d]) + 0.5) / (n - 1) for d in range(dim)]
g = np.cos(k * xs[0])[:, None] * np.cos(k * xs[1])[None, :] if dim == 2 else \
    np.cos(k * xs[0])[:, None, None] * np.cos(k * xs[1])[None, :, None] * np.cos(k * xs[2])[None, None, :]
pt.from_numpy(np.asfortranarray(g))
lap = (lambda f: d2x(D2, f) + d2y(D2, f)) if dim == 2 else (lambda f: d2x(D2, f) + d2y(D2, f) + d2z(D2, f))
bf.assign(lap(pt))
h = EqnSolveHandler(lambda e: (lap(e), bf), p, type_=ST.PCG, precond=ST.PFMG, tol=1e-10, maxIter=100, pinValue=True, staticMat=True)
p.assign(0.0)
st = h.solve()
ms = C.c_float()
tot = 0.0
for _ in range(args.solves):
    p.assign(0.0)
    capi.check(l.opf_synchronize())
    if world > 1:
        dist.barrier()
    capi.check(l.opf_timer_begin())
    st = h.solve()
    capi.check(l.opf_timer_end(C.byref(ms)))
    tot += ms.value
t = torch.tensor([tot / args.solves], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    cells = 1
    for d in range(dim):
        cells *= dims[d] - 1
    print(json.dumps({"config": f"Poisson {dim}-D, {'x'.join(str(x - 1) for x in dims)} cells, {'periodic' if dim == 3 else 'Neumann'} + pin, PCG + GMG, tol 1e-10",
                      "n_gpus": world, "ms_per_solve": t.item(), "iterations": st.niter, "relres": st.relerr, "levels": h.levels(),
                      "cells_x_iterations_per_s": cells * st.niter / (t.item() * 1e-3)}), flush=True)
if world > 1:
    capi.check(l.opf_comm_finalize())
    dist.destroy_process_group()
