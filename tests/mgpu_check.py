"""Multi-GPU parity check, launched as  python -m torch.distributed.run --nproc-per-node N tests/mgpu_check.py
One process per GPU; the field is slab-decomposed along its slowest axis (opf_split_slab), halos travel over NCCL
(engine_comm.cu), and every rank compares its block against a single-GPU run of the SAME global problem made in the same
process (a field without a split strategy) -- bit for bit, since both paths run the same device functors.
Reference behaviour being checked: updatePadding's MPI branch, CartesianField.hpp:630-768, and globalReduce,
RangeFor.hpp:125-135."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opflow_b200 import capi, host  # noqa: E402
from opflow_b200.host import D2SecondOrderCentered as D2, d2x, d2y, d2z  # noqa: E402


def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    l = capi.lib()
    capi.check(l.opf_init(lrank))
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = (C.c_ubyte * 128)()
        capi.check(l.opf_comm_unique_id(raw))
        idbuf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    dist.broadcast(idbuf, 0)
    raw = (C.c_ubyte * 128)(*idbuf.cpu().tolist())
    capi.check(l.opf_comm_init(rank, world, raw))
    failures = []

    def build(dims, periodic, pad, split, ext=None):
        mb = host.MeshBuilder(3).newMesh(*dims)
        for d in range(3):
            mb.setMeshOfDim(d, 0., 1. + d)
        mesh = mb.build()
        b = host.ExprBuilder().setName("u").setMesh(mesh)
        for d in range(3):
            if periodic:
                b.setBC(d, 0, host.BCType.Periodic).setBC(d, 1, host.BCType.Periodic)
            else:
                b.setBC(d, 0, host.BCType.Dirc, 1.).setBC(d, 1, host.BCType.Neum if d == 1 else host.BCType.Dirc, 0.5)
        b.setExt(pad if ext is None else ext).setPadding(pad)
        if split:
            b.setSplitStrategy(world, rank, host.split_slab(mesh, world))
        return mesh, b.build()

    for name, dims, periodic, pad, steps in (("dirichlet/neumann box", (70, 37, 16 * world + 1), False, 1, 12),
                                             ("periodic box pad 2", (66, 34, 12 * world + 1), True, 2, 9),
                                             ("wide rows (TMA skeleton)", (141, 21, 10 * world + 1), False, 1, 6),
                                             ("thick slabs (pipelined host route)", (97, 19, 40 * world + 1), False, 1, 3)):
        for mode in (capi.MODE_EXACT, capi.MODE_FAST):
            host.set_mode(mode)
            ext = 0 if name.startswith("thick") else None  # no ghost cells to refresh: the host route pipelines (bench.py's field)
            mesh_g, g = build(dims, periodic, pad, False, ext)
            mesh_s, s = build(dims, periodic, pad, True, ext)
            full = g.localRange
            rng = np.random.default_rng(7)
            init = rng.standard_normal(full.shape(3))
            g.from_numpy(init)
            lr = s.localRange
            sl = tuple(slice(lr.start[d] - full.start[d], lr.end[d] - full.start[d]) for d in range(3))
            s.from_numpy(init[sl])
            c = 0.05 * min((1. + d) / (dims[d] - 1) for d in range(3)) ** 2
            eg = g + c * (d2x(D2, g) + d2y(D2, g) + d2z(D2, g))
            es = s + c * (d2x(D2, s) + d2y(D2, s) + d2z(D2, s))
            for _ in range(steps):
                g.assign(eg)
                s.assign(es)
            a, r = s.to_numpy(), g.to_numpy()[sl]
            if not np.array_equal(a, r):
                failures.append(f"{name} mode={mode}: block differs, max abs {np.abs(a - r).max():.3e}")
            # ghost planes too (what the next stencil sweep would read): the exchanged halo planes over the block's own x-y extent
            # must hold the neighbour's values -- faces travel straight from / into the pitched storage (engine_comm.cu)
            glo = [lr.start[0], lr.start[1], max(full.start[2], lr.start[2] - pad)]
            ghi = [lr.end[0], lr.end[1], min(full.end[2], lr.end[2] + pad)]
            gh = s.to_numpy(capi.Range.make(glo, ghi))
            gsl = tuple(slice(glo[d] - full.start[d], ghi[d] - full.start[d]) for d in range(3))
            if not np.array_equal(gh, g.to_numpy()[gsl]):
                failures.append(f"{name} mode={mode}: halo planes differ from the single-GPU field")
            # host-buffer route (opf_assign_host) on the decomposed field: boundary chunks first, input halo exchange, chunked
            # sweeps and downloads -- must equal one more resident step, block and halo planes alike
            if True:
                import torch as _t
                hin = _t.empty(tuple(lr.shape(3))[::-1], dtype=_t.float64).pin_memory()
                hout = _t.empty_like(hin).pin_memory()
                cur = s.to_numpy()
                hin.numpy()[...] = np.ascontiguousarray(cur.transpose(2, 1, 0))
                sig, fields, scalars = es.flatten()
                F = (C.c_void_p * len(fields))(*[f.h for f in fields])
                S = (C.c_double * len(scalars))(*scalars)
                capi.check(l.opf_assign_host(s.h, capi.OP_EQ, sig.encode(), F, len(fields), S, len(scalars), s.h, C.c_void_p(hin.data_ptr()),
                                             C.c_void_p(hout.data_ptr())))
                g.assign(eg)
                r2 = g.to_numpy()
                if not np.array_equal(hout.numpy().transpose(2, 1, 0), r2[sl]):
                    failures.append(f"{name} mode={mode}: opf_assign_host result differs from the resident step")
                if not np.array_equal(s.to_numpy(capi.Range.make(glo, ghi)), r2[gsl]):
                    failures.append(f"{name} mode={mode}: halo planes after opf_assign_host differ")
            # globalReduce = local reduce + allreduce
            loc = host.rangeReduce(s, capi.RED_SUM)
            v = (C.c_double * 1)(loc)
            capi.check(l.opf_comm_allreduce(v, 1, capi.RED_SUM))
            ref = host.rangeReduce(g, capi.RED_SUM)
            if abs(v[0] - ref) > 1e-11 * max(1.0, abs(ref)):
                failures.append(f"{name} mode={mode}: global sum {v[0]!r} vs {ref!r}")
            del g, s, eg, es
    # ---- EvenSplitStrategy blocks (the reference's default strategy: 2 x 2 blocks in 2-D on 4 ranks, with edge / corner
    # neighbours -> the generic, non-overlapped exchange; an x-split on 2 ranks -> sub-boxes shifted along the fastest axis)
    for name, dims in (("even split 2-D", (70, 52)), ("even split 3-D", (38, 30, 26))):
        dim = len(dims)
        for mode in (capi.MODE_EXACT, capi.MODE_FAST):
            host.set_mode(mode)

            def mk2(split):
                mb = host.MeshBuilder(dim).newMesh(*dims)
                for d in range(dim):
                    mb.setMeshOfDim(d, 0., 1. + d)
                mesh = mb.build()
                b = host.ExprBuilder().setName("u").setMesh(mesh).setLoc([1] * dim).setExt(1).setPadding(1)
                for d in range(dim):
                    b.setBC(d, 0, host.BCType.Dirc, 1.).setBC(d, 1, host.BCType.Neum, 0.25)
                if split:
                    b.setSplitStrategy(world, rank, host.split_even(mesh, world))
                return b.build()

            g, sf = mk2(False), mk2(True)
            full, lr = g.localRange, sf.localRange
            init = np.random.default_rng(21).standard_normal(full.shape(dim))
            g.from_numpy(init)
            sl = tuple(slice(lr.start[d] - full.start[d], lr.end[d] - full.start[d]) for d in range(dim))
            sf.from_numpy(init[sl])
            c = 0.05 * min((1. + d) / (dims[d] - 1) for d in range(dim)) ** 2
            lapg = d2x(D2, g) + d2y(D2, g) if dim == 2 else d2x(D2, g) + d2y(D2, g) + d2z(D2, g)
            laps = d2x(D2, sf) + d2y(D2, sf) if dim == 2 else d2x(D2, sf) + d2y(D2, sf) + d2z(D2, sf)
            for _ in range(8):
                g.assign(g + c * lapg)
                sf.assign(sf + c * laps)
            a_, r_ = sf.to_numpy(), g.to_numpy()[sl]
            if not np.array_equal(a_, r_):
                failures.append(f"{name} mode={mode}: block differs, max abs {np.abs(a_ - r_).max():.3e} (neighbours: {len(sf.neighbors())})")
            elif rank == 0 and mode == capi.MODE_EXACT:
                print(f"  {name}: {len(sf.neighbors())} neighbours on rank 0, block {lr.tup(dim)}", flush=True)
            del g, sf
    # ---- resplitWithStrategy (CartesianField.hpp:83-177): z-slabs -> y-slabs -> z-slabs, values kept; the field must behave like one
    # BUILT with the new map (ranges, neighbours, halo exchange through the staged route) -- compared against the one-GPU field
    for mode in (capi.MODE_EXACT,):
        host.set_mode(mode)
        dims = (37, 12 * world + 1, 8 * world + 1)
        mesh_g, g = build(dims, False, 1, False)
        mesh_s, sf = build(dims, False, 1, True)
        full = g.localRange
        init = np.random.default_rng(5).standard_normal(full.shape(3))
        g.from_numpy(init)
        lr = sf.localRange
        sf.from_numpy(init[tuple(slice(lr.start[d] - full.start[d], lr.end[d] - full.start[d]) for d in range(3))])
        mr, _ = mesh_s.ranges()

        def blocks_along(ax):
            out = []
            n = mr.end[ax] - 1 - mr.start[ax]
            for r_ in range(world):
                st_, en_ = list(mr.start[:3]), [mr.end[d] - 1 for d in range(3)]
                st_[ax] = mr.start[ax] + (n * r_) // world
                en_[ax] = mr.start[ax] + (n * (r_ + 1)) // world
                out.append(capi.Range.make(st_, en_))
            return out

        c = 0.05 * min((1. + d) / (dims[d] - 1) for d in range(3)) ** 2
        for ax in (1, 2, 0):
            sf.resplit(blocks_along(ax))
            lr = sf.localRange
            sl = tuple(slice(lr.start[d] - full.start[d], lr.end[d] - full.start[d]) for d in range(3))
            want_lo = mr.start[ax] + ((mr.end[ax] - 1 - mr.start[ax]) * rank) // world
            if lr.start[ax] != want_lo or any(lr.start[d] != full.start[d] or lr.end[d] != full.end[d] for d in range(3) if d != ax):
                failures.append(f"resplit along {ax}: local range {lr.tup(3)}")
                continue
            if not np.array_equal(sf.to_numpy(), g.to_numpy()[sl]):
                failures.append(f"resplit along {ax}: values moved wrongly")
                continue
            for _ in range(4):
                g.assign(g + c * (d2x(D2, g) + d2y(D2, g) + d2z(D2, g)))
                sf.assign(sf + c * (d2x(D2, sf) + d2y(D2, sf) + d2z(D2, sf)))
            if not np.array_equal(sf.to_numpy(), g.to_numpy()[sl]):
                failures.append(f"resplit along {ax}: sweeps after the move differ from one GPU (neighbours: {len(sf.neighbors())})")
            elif rank == 0:
                print(f"  resplit along axis {ax}: block {lr.tup(3)}, {len(sf.neighbors())} neighbours, 4 sweeps bit-identical to one GPU", flush=True)
        del g, sf
    # ---- implicit path on a decomposed target: PCG + geometric multigrid with distributed levels (halo exchange per level, global
    # dot products / mean projections, replicated coarse hierarchy behind one allreduce) against the same solve on one GPU
    from opflow_b200.host import EqnSolveHandler, StructSolverType as ST
    host.set_mode(capi.MODE_FAST)
    for name, dims, bctype, loc, pin in (("poisson 2-D neumann+pin", (257, 64 * world + 1), host.BCType.Neum, [1, 1], True),
                                         ("poisson 3-D dirichlet", (65, 33, 16 * world + 1), host.BCType.Dirc, [0, 0, 0], False),
                                         ("poisson 3-D periodic+pin", (33, 33, 16 * world + 1), host.BCType.Periodic, [1, 1, 1], True)):
        dim = len(dims)

        def mk(nm, split):
            mb = host.MeshBuilder(dim).newMesh(*dims)
            for d in range(dim):
                mb.setMeshOfDim(d, 0., 1.)
            mesh = mb.build()
            b = host.ExprBuilder().setName(nm).setMesh(mesh).setLoc(loc).setExt(1)
            for d in range(dim):
                if bctype == host.BCType.Periodic:
                    b.setBC(d, 0, bctype).setBC(d, 1, bctype)
                else:
                    b.setBC(d, 0, bctype, 0.).setBC(d, 1, bctype, 0.)
            if split:
                b.setPadding(1).setSplitStrategy(world, rank, host.split_slab(mesh, world))
            return b.build()

        lap = (lambda f: d2x(D2, f) + d2y(D2, f)) if dim == 2 else (lambda f: d2x(D2, f) + d2y(D2, f) + d2z(D2, f))
        res = {}
        for split in (False, True):
            p_, b_, t_ = mk("p", split), mk("b", split), mk("pt", split)
            full = mk("full", False).localRange if split else p_.localRange
            lr = t_.localRange
            rng = np.random.default_rng(11)
            fa = [np.cos(np.pi * (1 + k) * np.linspace(0., 1., full.shape(dim)[k])) for k in range(dim)]
            g = fa[0][:, None] * fa[1][None, :] if dim == 2 else fa[0][:, None, None] * fa[1][None, :, None] * fa[2][None, None, :]
            sl = tuple(slice(lr.start[d] - full.start[d], lr.end[d] - full.start[d]) for d in range(dim))
            t_.from_numpy(np.asfortranarray(g[sl]))
            b_.assign(lap(t_))
            p_.assign(0.0)
            h = EqnSolveHandler(lambda e: (lap(e), b_), p_, type_=ST.PCG, precond=ST.PFMG, tol=1e-11, maxIter=60, pinValue=pin, staticMat=True)
            st = h.solve()
            res[split] = (p_.to_numpy(), st.niter, st.relerr, h.levels(), sl)
            del h
        (pg, ng, rg, lg, _), (ps, ns, rs, ls, sl) = res[False], res[True]
        ref = pg[sl]
        err = np.abs(ps - ref).max() / max(np.abs(ref).max(), 1e-300)
        if not (err <= 1e-8 and rs <= 1e-11 and ns <= ng + 4):
            failures.append(f"{name}: decomposed solve iters={ns} (1 GPU: {ng}) relres={rs:.2e} levels={ls}/{lg} solution diff={err:.2e}")
        elif rank == 0:
            print(f"  {name}: iters {ns} (1 GPU {ng}), levels {ls} (1 GPU {lg}), solution diff {err:.1e}", flush=True)
    flag = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flag)
    if failures:
        print(f"[rank {rank}] FAIL: " + "; ".join(failures), flush=True)
    if rank == 0:
        print("MGPU_CHECK " + ("OK" if flag.item() == 0 else "FAILED") + f" world={world} launches={l.opf_launch_count()}", flush=True)
    capi.check(l.opf_comm_finalize())
    dist.destroy_process_group()
    return 0 if flag.item() == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
