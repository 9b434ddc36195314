#!/usr/bin/env python
"""tools/bench_c5.py -- BASELINE config C5 in reduced form: a 3-D periodic Taylor-Green box on a MAC-staggered mesh, z-slabs
across the GPUs of one node (weak scaling: --nz cells of z per GPU), NCCL halo exchange, distributed multigrid.

    python -m torch.distributed.run --nproc-per-node N tools/bench_c5.py [--size 1025] [--nz 128] [--steps 5]

One step = explicit viscous predictor for (u, v, w), pressure Poisson solve d2x(e)+d2y(e)+d2z(e) == div(u*)/dt (periodic, pinned),
projection u = u* - dt*grad(dp): the operator set of examples/LidDriven/LidDriven3D.cpp:59-88 without the implicit convection
terms, on the IC of SURVEY 8d (u = sin x cos y cos z, v = -cos x sin y cos z, w = 0).  Reports explicit GLUPS (six
sweeps per step), Poisson ms per solve and max |div u| after the projection (a rank-boundary-crossing correctness property)."""
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
from opflow_b200.host import (D1FirstOrderCentered as D1, D2SecondOrderCentered as D2, EqnSolveHandler, StructSolverType as ST, d2x, d2y, d2z,
                              dx, dy, dz)  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=1025, help="nodes along x and y")
ap.add_argument("--nz", type=int, default=128, help="cells along z per GPU")
ap.add_argument("--steps", type=int, default=5)
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
n, nzc = args.size, args.nz * world
two_pi = 2 * np.pi
Lz = two_pi * nzc / (n - 1)  # same spacing on every axis: weak scaling stretches the box along z
mesh = host.MeshBuilder(3).newMesh(n, n, nzc + 1).setMeshOfDim(0, 0., two_pi).setMeshOfDim(1, 0., two_pi).setMeshOfDim(2, 0., Lz).build()
split = host.split_slab(mesh, world) if world > 1 else None


def mk(name, loc):
    b = host.ExprBuilder().setMesh(mesh).setName(name).setLoc(loc).setExt(1).setPadding(1)
    for d in range(3):
        b.setBC(d, 0, host.BCType.Periodic).setBC(d, 1, host.BCType.Periodic)
    if split:
        b.setSplitStrategy(world, rank, split)
    return b.build()


LU, LV, LW, LP = [0, 1, 1], [1, 0, 1], [1, 1, 0], [1, 1, 1]
u, v, w = mk("u", LU), mk("v", LV), mk("w", LW)
us, vs, ws = mk("us", LU), mk("vs", LV), mk("ws", LW)
dp, rhs, div = mk("dp", LP), mk("rhs", LP), mk("div", LP)


def coords(f, loc):
    lr = f.localRange
    h = two_pi / (n - 1), two_pi / (n - 1), two_pi / nzc  # z in units where the box is 2*pi periodic
    return [(np.arange(lr.start[d], lr.end[d]) + (0.5 if loc[d] else 0.0)) * h[d] for d in range(3)]


x, y, z = coords(u, LU)
u.from_numpy(np.asfortranarray(np.sin(x)[:, None, None] * np.cos(y)[None, :, None] * np.cos(z)[None, None, :]))
x, y, z = coords(v, LV)
v.from_numpy(np.asfortranarray(-np.cos(x)[:, None, None] * np.sin(y)[None, :, None] * np.cos(z)[None, None, :]))
w.assign(0.0)
dt, nu = 1e-3, 1e-2
lap = lambda f: d2x(D2, f) + d2y(D2, f) + d2z(D2, f)
pred = [(us, u + (dt * nu) * lap(u)), (vs, v + (dt * nu) * lap(v)), (ws, w + (dt * nu) * lap(w))]
divexpr = (dx(D1, us) + dy(D1, vs) + dz(D1, ws)) / dt
corr = [(u, us - dt * dx(D1, dp)), (v, vs - dt * dy(D1, dp)), (w, ws - dt * dz(D1, dp))]
h = EqnSolveHandler(lambda e: (lap(e), rhs), dp, type_=ST.PCG, precond=ST.PFMG, tol=1e-10, maxIter=100, pinValue=True, staticMat=True)


def sync():
    capi.check(l.opf_synchronize())
    if world > 1:
        dist.barrier()


ms = C.c_float()
t_exp = t_sol = 0.0
iters = 0
for step in range(args.steps + 1):  # step 0 is warm-up (allocations, multigrid set-up)
    sync()
    capi.check(l.opf_timer_begin())
    for dst, e in pred:
        dst.assign(e)
    rhs.assign(divexpr)
    capi.check(l.opf_timer_end(C.byref(ms)))
    a = ms.value
    dp.assign(0.0)
    sync()
    capi.check(l.opf_timer_begin())
    st = h.solve()
    capi.check(l.opf_timer_end(C.byref(ms)))
    b = ms.value
    sync()
    capi.check(l.opf_timer_begin())
    for dst, e in corr:
        dst.assign(e)
    capi.check(l.opf_timer_end(C.byref(ms)))
    if step > 0:
        t_exp += a + ms.value
        t_sol += b
        iters += st.niter
div.assign((dx(D1, u) + dy(D1, v) + dz(D1, w)) / 1.0)
dmax = host.rangeReduce(host.abs_(div), capi.RED_MAX)
umax = host.rangeReduce(host.abs_(u), capi.RED_MAX)
t = torch.tensor([t_exp / args.steps, t_sol / args.steps, dmax, umax], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    cells = (n - 1) * (n - 1) * nzc
    exp_ms, sol_ms = t[0].item(), t[1].item()
    print(json.dumps({"config": f"C5-lite TaylorGreen3D {n - 1}x{n - 1}x{nzc} cells, MAC staggering, periodic, z-slabs x{world}", "n_gpus": world,
                      "explicit_ms_per_step": exp_ms, "explicit_sweeps_per_step": 7, "explicit_glups": 7 * cells / (exp_ms * 1e-3) / 1e9,
                      "poisson_ms_per_solve": sol_ms, "poisson_iterations": iters / args.steps, "poisson_levels": h.levels(),
                      "max_abs_div_after_projection": t[2].item(), "max_abs_u": t[3].item()}), flush=True)
if world > 1:
    capi.check(l.opf_comm_finalize())
    dist.destroy_process_group()
