"""N>1 host logic on CPU: two gloo ranks plan a slab-decomposed field (opf_field_plan -- no device), exchange halos with
torch.distributed send/recv following the engine's neighbour lists and message order, and check every received ghost
plane against the global function.  Mirrors the reference's MPI tests (test/Core/Field/CartesianFieldMPITest.cpp:
updatePadding after a decomposed assignment) for the host half of engine_comm.cu."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opflow_b200 import host


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _global(idx, dims):
    """deterministic value of global node (i, j, k)"""
    i, j, k = idx
    return i + 1000.0 * j + 1e6 * k + 0.5


def _inverse_code(code, dim):
    out, p = 0, 1
    for _ in range(dim):
        d = (code // p) % 3
        d = 2 if d == 1 else (1 if d == 2 else 0)
        out += d * p
        p *= 3
    return out


def _worker(rank, world, port, case, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dims, periodic_axis, pad = case
        dim = len(dims)
        mb = host.MeshBuilder(dim).newMesh(*dims)
        for d in range(dim):
            mb.setMeshOfDim(d, 0., 1.)
        mesh = mb.build()
        b = host.ExprBuilder().setName("u").setMesh(mesh)
        for d in range(dim):
            if d == periodic_axis:
                b.setBC(d, 0, host.BCType.Periodic).setBC(d, 1, host.BCType.Periodic)
            else:
                b.setBC(d, 0, host.BCType.Dirc, 1.).setBC(d, 1, host.BCType.Dirc, 1.)
        b.setPadding(pad).setExt(pad).setSplitStrategy(world, rank, host.split_slab(mesh, world))
        u = b.plan()
        lr, st = u.localRange, u.storageRange
        nb = u.neighbors()
        # every rank's neighbour list, to check the pairing: what I send to r must be what r expects from me
        allnb = [None] * world
        dist.all_gather_object(allnb, (lr.tup(dim), nb))
        period = [dims[d] - 1 for d in range(dim)]
        for (peer, send, recv, code) in nb:
            peer_nb = allnb[peer][1]
            match = [x for x in peer_nb if x[0] == rank and x[3] == _inverse_code(code, dim)]
            assert len(match) == 1, (rank, peer, code, peer_nb)
            psend, precv = match[0][1], match[0][2]
            # the peer's send box is my recv box shifted by the periodic image
            ext_s = [psend[1][d] - psend[0][d] for d in range(dim)]
            ext_r = [recv[1][d] - recv[0][d] for d in range(dim)]
            assert ext_s == ext_r, (rank, peer, psend, recv)
            for d in range(dim):
                sh = recv[0][d] - psend[0][d]
                assert sh in (0, period[d], -period[d]), (rank, peer, d, sh)
        # simulate the exchange on numpy storage
        shape = [st.end[d] - st.start[d] for d in range(dim)]
        a = np.full(shape, np.nan)
        sl = tuple(slice(lr.start[d] - st.start[d], lr.end[d] - st.start[d]) for d in range(dim))
        grids = np.meshgrid(*[np.arange(lr.start[d], lr.end[d]) for d in range(dim)], indexing="ij")
        gi = [grids[d] if d < dim else 0 for d in range(3)]
        a[sl] = _global(gi, dims)

        def box(r):
            return tuple(slice(r[0][d] - st.start[d], r[1][d] - st.start[d]) for d in range(dim))

        sorder = sorted(range(len(nb)), key=lambda i: nb[i][3])
        rorder = sorted(range(len(nb)), key=lambda i: _inverse_code(nb[i][3], dim))
        reqs, bufs = [], []
        for i in sorder:
            t = torch.from_numpy(np.ascontiguousarray(a[box(nb[i][1])]))
            reqs.append(dist.isend(t, nb[i][0]))
            bufs.append(t)
        recvs = []
        for i in rorder:
            r = nb[i][2]
            if any(r[1][d] <= r[0][d] for d in range(dim)):
                continue
            t = torch.empty([r[1][d] - r[0][d] for d in range(dim)], dtype=torch.float64)
            reqs.append(dist.irecv(t, nb[i][0]))
            recvs.append((i, t))
        for rq in reqs:
            rq.wait()
        for i, t in recvs:
            a[box(nb[i][2])] = t.numpy()
        # every ghost plane that belongs to another rank's block (or its periodic image) now holds the global function
        checked = 0
        for (peer, send, recv, code) in nb:
            grids = np.meshgrid(*[np.arange(recv[0][d], recv[1][d]) for d in range(dim)], indexing="ij")
            gi = []
            for d in range(dim):
                g = grids[d].copy()
                if d == periodic_axis:
                    g = np.where(g < 0, g + period[d], g)
                    g = np.where(g >= period[d], g - period[d], g)
                gi.append(g)
            while len(gi) < 3:
                gi.append(0)
            want = _global(gi, dims)
            got = a[box(recv)]
            assert np.array_equal(got, want), (rank, peer, code)
            checked += got.size
        q.put((rank, checked, len(nb)))
    finally:
        dist.destroy_process_group()


CASES = {
    "slab3d_dirichlet_pad1": ((17, 13, 33), None, 1),
    "slab3d_periodic_z_pad2": ((9, 9, 33), 2, 2),
    "slab2d_periodic_y_pad2": ((17, 33), 1, 2),
    "slab2d_dirichlet_pad3": ((21, 41), None, 3),
}


@pytest.mark.parametrize("name", list(CASES))
def test_slab_halo_plan_two_ranks(name):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, CASES[name], q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    res = sorted(q.get(timeout=5) for _ in range(world))
    for rank, checked, nnb in res:
        assert nnb >= 1 and checked > 0
    if CASES[name][1] is not None:  # periodic: both ranks see two neighbours (the seam links the last slab to the first)
        assert all(nnb == 2 for _, _, nnb in res)


def _resplit_worker(rank, world, port, case, q):
    """resplitWithStrategy's host half (opf_field_resplit_plan) on plan-only fields: every rank moves its block of the global function
    to the new decomposition with gloo send/recv following the engine's send / recv boxes"""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dims, loc, new_axis = case
        dim = len(dims)
        mb = host.MeshBuilder(dim).newMesh(*dims)
        for d in range(dim):
            mb.setMeshOfDim(d, 0., 1.)
        mesh = mb.build()
        b = host.ExprBuilder().setName("u").setMesh(mesh).setLoc(loc)
        for d in range(dim):
            b.setBC(d, 0, host.BCType.Dirc, 1.).setBC(d, 1, host.BCType.Neum, 0.)
        b.setPadding(1).setExt(1).setSplitStrategy(world, rank, host.split_slab(mesh, world))
        u = b.plan()
        lr = u.localRange
        mr, _ = mesh.ranges()
        ncell = mr.end[new_axis] - 1 - mr.start[new_axis]
        new_map = []
        for r_ in range(world):
            st_, en_ = [mr.start[d] for d in range(3)], [mr.end[d] - 1 if d < dim else 1 for d in range(3)]
            st_[new_axis] = mr.start[new_axis] + (ncell * r_) // world
            en_[new_axis] = mr.start[new_axis] + (ncell * (r_ + 1)) // world
            new_map.append(host.Range.make(st_, en_))
        send, recv, nl = u.resplit_plan(new_map)

        def glob(r):
            grids = np.meshgrid(*[np.arange(r.start[d], r.end[d]) for d in range(dim)], indexing="ij")
            gi = [grids[d] if d < dim else 0 for d in range(3)]
            return _global(gi, dims)

        def cnt(r):
            return int(np.prod([max(0, r.end[d] - r.start[d]) for d in range(dim)]))

        # what I send must be what the peer expects from me (box for box), and my send boxes partition my old block
        allb = [None] * world
        dist.all_gather_object(allb, ([s.tup(dim) for s in send], [r.tup(dim) for r in recv], lr.tup(dim), nl.tup(dim)))
        for peer in range(world):
            peer_expects = allb[peer][1][rank]
            peer_cnt = int(np.prod([max(0, e - s) for s, e in zip(*peer_expects)]))
            assert (cnt(send[peer]) == 0 and peer_cnt == 0) or send[peer].tup(dim) == peer_expects, (rank, peer)
        assert sum(cnt(s) for s in send) == cnt(lr), "the send boxes do not cover the old block exactly once"
        assert sum(cnt(r) for r in recv) == cnt(nl), "the recv boxes do not cover the new block exactly once"
        old = glob(lr)
        new = np.full([nl.end[d] - nl.start[d] for d in range(dim)], np.nan)
        reqs, keep, got = [], [], []
        for peer in range(world):
            if cnt(send[peer]) and peer != rank:
                sl = tuple(slice(send[peer].start[d] - lr.start[d], send[peer].end[d] - lr.start[d]) for d in range(dim))
                t = torch.from_numpy(np.ascontiguousarray(old[sl]))
                keep.append(t)
                reqs.append(dist.isend(t, peer))
            if cnt(recv[peer]) and peer != rank:
                t = torch.empty([recv[peer].end[d] - recv[peer].start[d] for d in range(dim)], dtype=torch.float64)
                reqs.append(dist.irecv(t, peer))
                got.append((peer, t))
        for rq in reqs:
            rq.wait()
        if cnt(send[rank]):
            s_, r_ = send[rank], recv[rank]
            assert s_.tup(dim) == r_.tup(dim)
            new[tuple(slice(r_.start[d] - nl.start[d], r_.end[d] - nl.start[d]) for d in range(dim))] = \
                old[tuple(slice(s_.start[d] - lr.start[d], s_.end[d] - lr.start[d]) for d in range(dim))]
        for peer, t in got:
            new[tuple(slice(recv[peer].start[d] - nl.start[d], recv[peer].end[d] - nl.start[d]) for d in range(dim))] = t.numpy()
        assert np.array_equal(new, glob(nl)), "moved block differs from the global function"
        # the new local range is what a field BUILT with the new map gets
        b2 = host.ExprBuilder().setName("u2").setMesh(mesh).setLoc(loc)
        for d in range(dim):
            b2.setBC(d, 0, host.BCType.Dirc, 1.).setBC(d, 1, host.BCType.Neum, 0.)
        b2.setPadding(1).setExt(1).setSplitStrategy(world, rank, new_map)
        assert b2.plan().localRange.tup(dim) == nl.tup(dim)
        q.put((rank, cnt(nl), sum(1 for p_ in range(world) if p_ != rank and cnt(send[p_]))))
    finally:
        dist.destroy_process_group()


RESPLIT_CASES = {
    "3d_corner_z_to_y": ((13, 17, 21), [0, 0, 0], 1),
    "3d_center_z_to_x": ((13, 9, 17), [1, 1, 1], 0),
    "2d_mac_y_to_x": ((21, 17), [0, 1], 0),
}


@pytest.mark.parametrize("name", list(RESPLIT_CASES))
@pytest.mark.parametrize("world", [2, 3])
def test_resplit_plan_moves_every_cell_once(name, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_resplit_worker, args=(r, world, port, RESPLIT_CASES[name], q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert all(n > 0 for _, n, _ in res)
    assert all(peers == world - 1 for _, _, peers in res)  # a slab -> cross-slab move talks to every other rank
