#!/usr/bin/env python
"""bench.py -- benchmarks of the B200 evaluation engine for OpFlow's stencil hot path, one JSON line per run (rank 0).

    python bench.py [--config C2] [--gpus N] [--steps K] [--warmup W] [--mode fast|exact] [--impl reference]

Default workload = BASELINE.json configs[1] (C2): 3-D FTCS heat equation, 513^3 nodes (511^3 interior updates per step), FP64, 7-point
stencil  u = u + dt*alpha*(d2x(u) + d2y(u) + d2z(u))  with D2SecondOrderCentered, Dirichlet 1 on every face -- the 3-D extension of
examples/FTCS2D/FTCS-OMP.cpp:26.  A "step" is one such assignment including the reference's implied updatePadding().
The other BASELINE configs run with --config C1 | C3 | C3B | C4 | C5; at N = 1 the default line also carries a compact "configs" block
(C1, C3, C4: value, roofline, cpu_baseline, e2e each) so that one driver run records every single-GPU configuration.

value     metric of the config with the operands resident in HBM; a run times `--batches` batches (default 5) of exactly K steps, each
          bracketed by barrier + synchronize, CUDA events on the engine stream, max over ranks; the MEDIAN batch is reported and the
          list is kept in "batches_ms" (one stalled batch -- nvidia-smi sampling, a first NCCL connection -- cannot decide the line).
e2e       the same metric through the C ABI with HOST buffers (pinned): H2D of the input and D2H of the result inside the timed region.
roofline  algorithmic bytes (SURVEY 8d per-update figure x updates of one launch) / mean launch duration of the dominant kernel, against
          MEASURED_PEAKS.json hbm_gbs; the kernel's name is what the engine reports it launched (opf_last_kernel_name).
cpu_baseline / --impl reference: the UNMODIFIED reference (oracle/_ref, TBB + HYPRE) on this box's host cores, bounded sample.
N > 1 (torchrun): weak scaling, z-slabs per GPU, NCCL halo exchange overlapped with the interior sweep; "parity_ok" is a
reduced-size replay of the same decomposed assignment against a non-decomposed field on every rank (bit-exact, halo planes included).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid-point updates/sec (GLUPS)"


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # under load = the upper half of the samples (the sampler also sees the idle gaps between the legs of the run)
        load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------ the reference on the host cores
def ref_env():
    from oracle import oracle as O
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(O.HERE, "_ref", "lib") + ":" + env.get("LD_LIBRARY_PATH", "")
    return env


def ref_explicit(case, n, steps, warmup, threads, timeout=1500):
    """oracle/_ref/bin/ref_explicit: unmodified OpFlow + TBB, the explicit configs (FTCS2D / FTCS3D / WENO5)"""
    from oracle import oracle as O
    if not O.ref_available("ref_explicit"):
        return None
    return O.run_ref("ref_explicit", "--case", case, "--n", n, "--steps", steps, "--warmup", warmup, "--threads", threads,
                     "--init", "sin", env=ref_env(), timeout=timeout)[0]


def ref_poisson(n, solves, threads, timeout=1500):
    """oracle/_ref/bin/ref_poisson: LidDriven2D's pressure handler (GMRES + PFMG through HYPRE) on (n-1)^2 cells"""
    from oracle import oracle as O
    if not O.ref_available("ref_poisson"):
        return None
    return O.run_ref("ref_poisson", "--n", n, "--solves", solves, "--threads", threads, env=ref_env(), timeout=timeout)[0]


def cpu_explicit(case, n, steps, what):
    cores = host_cores()
    try:
        t0 = time.time()
        r = ref_explicit(case, n, steps, 2, cores)
        if r:
            return {"value": r["mlups"] / 1e3, "unit": "GLUPS", "cores": cores, "kind": "reference",
                    "sample": f"reference {what}, {steps} steps after 2 warm-up, {cores} TBB threads, wall {time.time() - t0:.1f}s"}
    except Exception as e:
        return {"value": None, "unit": "GLUPS", "cores": cores, "kind": "reference", "sample": f"failed: {e}"[:160]}
    return None


def workload_c2(n, nz_nodes):
    return f"FTCS3D heat equation {n}x{n}x{nz_nodes} nodes FP64 7-point explicit, u = u + dt*alpha*(d2x+d2y+d2z)(u), Dirichlet 1"


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the selected config on this box's host cores."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    cores = host_cores()
    cfg = args.config
    t0 = time.time()
    base = {"impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
    try:
        if cfg == "C4":
            n = 1025  # bounded sample: one solve of the reference takes ~5 s at 1025^2 and ~90 s at 4097^2
            solves = max(2, min(args.steps, 3))
            r = ref_poisson(n, solves, cores)
            if r is None:
                raise RuntimeError("oracle/_ref/bin/ref_poisson missing (run oracle/build_ref.sh)")
            cells = (n - 1) ** 2
            val = cells * r["niter"] / (r["ms_per_solve"] * 1e-3) / 1e9
            sample = (f"LidDriven2D pressure Poisson on {n - 1}^2 cells (bounded sample of the 4096^2 workload), {solves} solves, GMRES+PFMG (HYPRE 2.33), "
                      f"{cores} threads, {r['ms_per_solve']:.0f} ms per solve, {r['niter']} outer iterations, wall {time.time() - t0:.1f}s")
            line = dict(base, metric="Poisson solve ms", value=r["ms_per_solve"] * (4096.0 / (n - 1)) ** 2, unit="ms", higher_is_better=False,
                        ms_per_step=r["ms_per_solve"], config={"workload": "C4 pressure Poisson 4096^2 cells, Neumann + pin, tol 1e-10 (reference timed at 1024^2 cells; "
                                                               "value scaled by the cell ratio 16)", "mode": "reference CPU path (HYPRE GMRES + PFMG)", "parallelism": "host cores only"},
                        cell_iterations_per_s=val * 1e9)
        else:
            case, full_n, cells_of = {"C1": ("ftcs2d", 1025, lambda m: (m - 2) ** 2), "C2": ("ftcs3d", args.n, lambda m: (m - 2) ** 3),
                                      "C3": ("weno_down", 2 ** 26 + 1, lambda m: m - 2), "C3B": ("weno_down", 2 ** 26 + 1, lambda m: m - 2),
                                      "C5": ("ftcs3d", args.n, lambda m: (m - 2) ** 3)}[cfg]
            # bounded sample: about two minutes of host work for K + W steps; the metric is a rate, the expression and arithmetic are the same
            rate = {"ftcs2d": 0.02e9, "ftcs3d": 0.012e9, "weno_down": 0.004e9}[case] * cores
            budget_cells = 120.0 * rate / max(1, args.steps + args.warmup)
            n = full_n
            if cells_of(n) > budget_cells:
                if case == "ftcs3d":
                    n = max(129, min(n, int(round(budget_cells ** (1.0 / 3.0))) + 2))
                elif case == "ftcs2d":
                    n = max(257, min(n, int(round(budget_cells ** 0.5)) + 2))
                else:
                    n = max(2 ** 16 + 1, min(n, 2 ** int(budget_cells).bit_length() // 2 + 1))
            r = None
            try:
                r = ref_explicit(case, n, args.steps, args.warmup, cores)
            except Exception as e:  # e.g. not enough host memory for 3 x 2 GiB: fall back to a smaller sample
                sys.stderr.write(f"reference at n={n} failed ({e}); retrying smaller\n")
                n = {"ftcs3d": 257, "ftcs2d": 513}.get(case, 2 ** 20 + 1)
                r = ref_explicit(case, n, args.steps, args.warmup, cores)
            if r is None:
                raise RuntimeError("oracle/_ref/bin/ref_explicit missing (run oracle/build_ref.sh)")
            glups = r["mlups"] / 1e3
            what = {"C1": "FTCS2D heat equation 1025x1025 nodes FP64 5-point explicit (examples/FTCS2D/FTCS-OMP.cpp:26)", "C2": workload_c2(args.n, args.n),
                    "C3": "CONV1D WENO5 advection, 2^26+1 nodes FP64, u = u - dt*c*dx<D1WENO53Downwind>(u)", "C3B": "CONV1D WENO5 advection, 2^26+1 nodes FP64",
                    "C5": workload_c2(args.n, args.n)}[cfg]
            sample = (f"{case} n={n}" + ("" if n == full_n else f" (bounded sample of the n={full_n} workload)")
                      + f", {args.steps} steps after {args.warmup} warm-up, {cores} TBB threads, wall {time.time() - t0:.1f}s")
            line = dict(base, metric=METRIC, value=glups, unit="GLUPS", ms_per_step=r["seconds"] / max(1, args.steps) * 1e3,
                        config={"workload": what, "mode": "reference CPU path (unmodified OpFlow, TBB rangeFor, all host cores)", "parallelism": "host cores only"})
    except Exception as e:
        print(json.dumps({"impl": "reference", "unavailable": f"{e}"[:200]}))
        return 0
    line["cpu_baseline"] = {"value": line["value"], "unit": line["unit"], "cores": cores, "kind": "reference", "sample": sample}
    line["e2e"] = {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ GPU side
class Env:
    """process-wide state of one bench run: engine, ranks, timing helpers"""

    def __init__(self, args):
        import torch
        from opflow_b200 import capi
        self.args, self.torch, self.capi = args, torch, capi
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: opflow_b200 has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.l = capi.lib()
        capi.check(self.l.opf_init(self.local_rank))
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if self.rank == 0:
                raw = (C.c_ubyte * 128)()
                capi.check(self.l.opf_comm_unique_id(raw))
                idbuf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
            dist.broadcast(idbuf, 0)
            raw = (C.c_ubyte * 128)(*idbuf.cpu().tolist())
            capi.check(self.l.opf_comm_init(self.rank, self.world, raw))
        self.peak, self.peak_src = peaks()

    def barrier(self):
        self.capi.check(self.l.opf_synchronize())
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, fn, steps, batches, barrier=True):
        """`batches` batches of exactly `steps` calls; CUDA events on the engine stream; -> list of batch times in ms (this rank)"""
        ms, out = C.c_float(), []
        for _ in range(batches):
            if barrier:
                self.barrier()
            else:
                self.capi.check(self.l.opf_synchronize())
            self.capi.check(self.l.opf_timer_begin())
            for _ in range(steps):
                fn()
            self.capi.check(self.l.opf_timer_end(C.byref(ms)))
            out.append(float(ms.value))
        if barrier:
            self.barrier()
        return out

    def max_over_ranks(self, values):
        if not self.dist:
            return list(values)
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def sum_over_ranks(self, values):
        if not self.dist:
            return list(values)
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.tolist()

    def kernel_name(self):
        return self.l.opf_last_kernel_name().decode()

    def roofline(self, bytes_per_update, updates, kernel_ms, kernel, traffic=None, note=None):
        achieved = bytes_per_update * updates / (kernel_ms * 1e-3) / 1e9
        r = {"bound": "hbm", "achieved": achieved, "peak": self.peak, "unit": "GB/s", "frac": achieved / self.peak, "traffic": traffic,
             "kernel": kernel, "kernel_ms": kernel_ms, "peak_source": self.peak_src, "bytes_per_update": bytes_per_update}
        if note:
            r["note"] = note
        return r

    def flusher(self):
        """a 512 MB device buffer rewritten between timed launches: evicts a field that would otherwise stay in the 126 MB L2"""
        if not hasattr(self, "_flush"):
            self._flush = self.torch.empty(64 * 1024 * 1024, dtype=self.torch.float64, device="cuda")
        stream = self.torch.cuda.ExternalStream(self.l.opf_stream())

        def flush():
            with self.torch.cuda.stream(stream):
                self._flush.add_(1.0)
        return flush


def flat(expr):
    sig, fields, scalars = expr.flatten()
    F = (C.c_void_p * max(1, len(fields)))(*[f.h for f in fields])
    S = (C.c_double * max(1, len(scalars)))(*scalars)
    return sig.encode(), F, len(fields), S, len(scalars)


def kernel_only(env, u, expr, steps, batches=3):
    """the dominant kernel alone: the same assignment without the trailing updatePadding() (OPF_ASSIGN_NO_PADDING), i.e. exactly one
    skeleton launch per iteration (ping-pong buffers as in a real step), launches back to back; median of `batches` means"""
    l, capi = env.l, env.capi
    sig, F, nf, S, ns = flat(expr)

    def one():
        capi.check(l.opf_assign_ex(u.h, capi.OP_EQ, sig, F, nf, S, ns, 1))
    for _ in range(3):
        one()
    name = env.kernel_name()
    t = env.timed(one, steps, batches, barrier=False)
    u.updatePadding()
    return statistics.median(t) / steps, name


def host_pinned(env, shape_f):
    """pinned host tensor holding an axis-0-fastest array of Fortran shape `shape_f`"""
    return env.torch.empty(tuple(shape_f)[::-1], dtype=env.torch.float64).pin_memory()


def e2e_assign(env, u, expr, init, steps, batches=3):
    """host buffers through opf_assign_host: H2D of the input field + the step + D2H of the result, every step"""
    import numpy as np
    l, capi = env.l, env.capi
    sig, F, nf, S, ns = flat(expr)
    shape = u.localRange.shape(u.dim)
    hin, hout = host_pinned(env, shape), host_pinned(env, shape)
    hin.numpy()[...] = np.ascontiguousarray(init.transpose(*range(u.dim - 1, -1, -1)))

    def step():
        capi.check(l.opf_assign_host(u.h, capi.OP_EQ, sig, F, nf, S, ns, u.h, C.c_void_p(hin.data_ptr()), C.c_void_p(hout.data_ptr())))
    for _ in range(2):
        step()
    t = env.timed(step, steps, batches)
    return t, hin.numel() * 8


# ---- C2 (default): FTCS3D, weak scaling in z-slabs
def build_ftcs3d(env, n, nz_cells, decomposed, lz):
    from opflow_b200 import host
    mesh = host.MeshBuilder(3).newMesh(n, n, nz_cells + 1).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).setMeshOfDim(2, 0., lz).build()
    b = host.ExprBuilder().setName("u").setMesh(mesh)
    for d in range(3):
        b.setBC(d, host.DimPos.start, host.BCType.Dirc, 1.).setBC(d, host.DimPos.end, host.BCType.Dirc, 1.)
    if decomposed:
        b.setPadding(1).setSplitStrategy(env.world, env.rank, host.split_slab(mesh, env.world))
    return b.build()


def parity_replay(env):
    """N > 1 self-check (the GPU-test box has one GPU): the decomposed assignment of this bench at reduced size -- 161 x 97 nodes in x-y,
    48 z-cells per rank -- against a NON-decomposed field of the same global problem held by every rank, 4 steps, both arithmetic modes:
    the block AND the exchanged halo planes must be bit-identical.  x-extent >= 64: the TMA skeleton runs, as in the timed loop."""
    import numpy as np
    from opflow_b200 import capi, host
    from opflow_b200.host import D2SecondOrderCentered as D2, d2x, d2y, d2z
    l = env.l
    ok, kern = True, ""
    nx, ny, nzc = 161, 97, 48 * env.world
    saved = l.opf_get_mode()
    for mode in (capi.MODE_EXACT, capi.MODE_FAST):
        host.set_mode(mode)
        mesh = host.MeshBuilder(3).newMesh(nx, ny, nzc + 1).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).setMeshOfDim(2, 0., float(env.world)).build()

        def mk(split):
            b = host.ExprBuilder().setName("u").setMesh(mesh)
            for d in range(3):
                b.setBC(d, 0, host.BCType.Dirc, 1.).setBC(d, 1, host.BCType.Dirc, 1.)
            b.setPadding(1)
            if split:
                b.setSplitStrategy(env.world, env.rank, host.split_slab(mesh, env.world))
            return b.build()
        g, s = mk(False), mk(True)
        full, lr = g.localRange, s.localRange
        init = np.asfortranarray(np.random.default_rng(2024).standard_normal(full.shape(3)))
        sl = tuple(slice(lr.start[d] - full.start[d], lr.end[d] - full.start[d]) for d in range(3))
        g.from_numpy(init)
        s.from_numpy(np.asfortranarray(init[sl]))
        c = 0.1 / (nx - 1) ** 2
        eg, es = g + c * (d2x(D2, g) + d2y(D2, g) + d2z(D2, g)), s + c * (d2x(D2, s) + d2y(D2, s) + d2z(D2, s))
        for _ in range(4):
            g.assign(eg)
            s.assign(es)
        kern = env.kernel_name()
        ref = g.to_numpy()
        glo = [lr.start[0], lr.start[1], max(full.start[2], lr.start[2] - 1)]
        ghi = [lr.end[0], lr.end[1], min(full.end[2], lr.end[2] + 1)]
        gsl = tuple(slice(glo[d] - full.start[d], ghi[d] - full.start[d]) for d in range(3))
        ok = ok and bool(np.array_equal(s.to_numpy(), ref[sl])) and bool(np.array_equal(s.to_numpy(capi.Range.make(glo, ghi)), ref[gsl]))
        del g, s, eg, es
    host.set_mode(saved)
    bad = env.sum_over_ranks([0.0 if ok else 1.0])[0]
    return bad == 0, kern


def nccl_log_tail():
    """transport lines of the engine's communicator when NCCL_DEBUG_FILE points at a file bench.py chose (see main)"""
    path = os.environ.get("OPF_BENCH_NCCL_LOG")
    if not path:
        return None
    out = []
    try:
        for ln in open(path.replace("%p", str(os.getpid())).replace("%h", os.uname().nodename), errors="replace"):
            if " via " in ln or "nranks" in ln or "NVLS" in ln or "isAllDirectP2p" in ln:
                out.append(ln.strip()[-160:])
    except Exception:
        return None
    # de-duplicate channel lines: keep the transport summary
    seen, keep = set(), []
    for ln in out:
        key = ln.split("NCCL INFO")[-1].strip()
        key = key.split("Channel")[0] + key.split(" via ")[-1] if " via " in key else key
        if key not in seen:
            seen.add(key)
            keep.append(ln)
    return keep[-12:]


def run_c2(env, args):
    import numpy as np
    from opflow_b200 import capi, host
    from opflow_b200.host import D2SecondOrderCentered as D2, d2x, d2y, d2z
    l, world, rank = env.l, env.world, env.rank
    n = args.n
    nz_cells = (n - 1) * world  # weak scaling: 512 z-cells per GPU
    u = build_ftcs3d(env, n, nz_cells, world > 1, float(world))
    u.assign(0.0)
    # non-trivial initial condition (deterministic; same on every run): product of sines, evaluated with numpy on the host
    lr = u.localRange
    xs = [np.linspace(0.0, 1.0 if d < 2 else float(world), (n if d < 2 else nz_cells + 1))[lr.start[d]:lr.end[d]] for d in range(3)]
    init = np.asfortranarray(np.sin(np.pi * xs[0])[:, None, None] * np.sin(np.pi * xs[1])[None, :, None] * np.sin(np.pi * xs[2])[None, None, :])
    u.from_numpy(init)
    c = 0.1 / (n - 1) ** 2 * 1.0
    expr = u + c * (d2x(D2, u) + d2y(D2, u) + d2z(D2, u))
    w = [min(u.assignableRange.end[d], lr.end[d]) - max(u.assignableRange.start[d], lr.start[d]) for d in range(3)]
    updates_rank = w[0] * w[1] * w[2]

    parity_ok, parity_kernel = (None, None)
    if world > 1:
        parity_ok, parity_kernel = parity_replay(env)

    def step():
        u.assign(expr)
    for _ in range(args.warmup):
        step()
    env.barrier()
    sampler = ClockSampler(env.local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(1.0)  # the sampler's start-up (a driver query per GPU) stays out of the first timed batch
    env.barrier()
    for _ in range(args.steps):  # one untimed batch with the sampler running: its first driver queries stall launches for tens of ms
        step()
    env.barrier()
    launches0 = l.opf_launch_count()
    step()
    launches_per_step = l.opf_launch_count() - launches0
    batches_ms = env.timed(step, args.steps, args.batches)
    step_kernel = env.kernel_name()
    kernel_ms, kernel = kernel_only(env, u, expr, args.steps)
    e2e_ms, nbytes = e2e_assign(env, u, expr, init, max(1, args.e2e_steps))
    clocks = sampler.stop() if rank == 0 else None

    batches_ms = env.max_over_ranks(batches_ms)
    e2e_ms = env.max_over_ranks(e2e_ms)
    kernel_ms = env.max_over_ranks([kernel_ms])[0]
    updates, launches_step_total = env.sum_over_ranks([float(updates_rank), float(launches_per_step)])
    if rank != 0:
        return None
    ms_per_step = statistics.median(batches_ms) / args.steps
    glups = updates / (ms_per_step * 1e-3) / 1e9
    e2e_step = statistics.median(e2e_ms) / max(1, args.e2e_steps)
    roof = env.roofline(16.0, updates_rank, kernel_ms,
                        f"{kernel}<Add<F0,Mul<S0,Add<Add<D2C<0,F1>,D2C<1,F2>>,D2C<2,F3>>>>, {args.mode}, alias0, CX=2, 64x4 threads, 16-plane march, ring 6>",
                        note="achieved = 16 B x updates of one launch / mean launch duration (CUDA events on the engine stream, launches back to back, "
                             "no BC / halo launches between them); whole_step_* adds the BC-face fill (and halo) launches")
    step_achieved = 16.0 * updates_rank / (ms_per_step * 1e-3) / 1e9
    roof["whole_step_achieved"], roof["whole_step_frac"], roof["step_kernel"] = step_achieved, step_achieved / env.peak, step_kernel
    try:  # DRAM bytes of one launch from the committed ncu capture of this kernel on this workload (a static figure, not re-measured here)
        prof = json.load(open(os.path.join(ROOT, "profiles", "ftcs3d_traffic.json")))
        if kernel == "opf::tma_kernel" and n == 513:
            roof["traffic"] = prof.get("dram_bytes_per_launch")
            roof["traffic_source"] = "static, from profiles/ftcs3d_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"
    except Exception:
        pass
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_explicit("ftcs3d", 257, 100, "ftcs3d 257^3 nodes")
    line = {"metric": METRIC, "value": glups, "unit": "GLUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_c2(n, nz_cells + 1), "mode": args.mode,
                       "l2": "field (1.1 GB per GPU) is larger than L2 (126 MB): no flush needed",
                       "parallelism": f"z-slabs x{world}, NCCL halo exchange overlapped with the interior sweep" if world > 1 else "single GPU"},
            "batches_ms": batches_ms, "timing": f"median of {args.batches} batches of {args.steps} steps, each bracketed by barrier + synchronize, max over ranks",
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": updates / (e2e_step * 1e-3) / 1e9, "unit": "GLUPS", "h2d_bytes_per_step": nbytes * world, "d2h_bytes_per_step": nbytes * world,
                    "steps": max(1, args.e2e_steps), "ms_per_step": e2e_step, "batches_ms": e2e_ms,
                    "api": "opf_assign_host: pinned host buffers, upload | sweep | download pipelined in z-chunks" + (", boundary chunks first + input halo exchange" if world > 1 else "")},
            "gpu_launches": int(round(launches_step_total * args.steps)), "clocks": clocks}
    if world > 1:
        line["parity_ok"] = bool(parity_ok)
        line["parity"] = {"what": "decomposed vs non-decomposed replay at 161x97x(48N+1) nodes, 4 steps, EXACT and FAST, block + halo planes bit-identical on every rank",
                          "kernel": parity_kernel}
        tail = nccl_log_tail()
        if tail is not None:
            line["comm_log_tail"] = tail
    return line


# ---- C1: FTCS2D 1025^2
def run_c1(env, args, compact=False):
    import numpy as np
    from opflow_b200 import host
    from opflow_b200.host import D2SecondOrderCentered as D2, d2x, d2y
    n = 1025
    mesh = host.MeshBuilder(2).newMesh(n, n).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).build()
    b = host.ExprBuilder().setName("u").setMesh(mesh)
    for d in range(2):
        b.setBC(d, 0, host.BCType.Dirc, 1.).setBC(d, 1, host.BCType.Dirc, 1.)
    u = b.build()
    u.assign(0.0)
    lr = u.localRange
    x = np.linspace(0., 1., n)
    init = np.asfortranarray(np.sin(np.pi * x)[:, None] * np.sin(np.pi * x)[None, :])
    u.from_numpy(init)
    expr = u + (0.1 / (n - 1) ** 2) * (d2x(D2, u) + d2y(D2, u))
    updates = (n - 2) ** 2
    steps = 400 if compact else max(args.steps, 200)

    def step():
        u.assign(expr)
    for _ in range(20):
        step()
    t = env.timed(step, steps, 3 if compact else args.batches, barrier=False)
    ms = statistics.median(t) / steps
    kern = env.kernel_name()
    kernel_ms, kname = kernel_only(env, u, expr, steps)
    # the same time loop through opf_assign_repeat: 32 steps per CUDA-graph launch, no gaps between the kernels
    sig, F, nf, S, ns = flat(expr)

    def replay():
        env.capi.check(env.l.opf_assign_repeat(u.h, env.capi.OP_EQ, sig, F, nf, S, ns, 34 + 32 * 30))
    replay()
    tr = env.timed(replay, 1, 3, barrier=False)
    replay_ms = statistics.median(tr) / (34 + 32 * 30)
    # the same step with the 8.4 MB field evicted from L2 before every launch (per-launch events: includes ~2 us of event overhead)
    flush = env.flusher()
    ev = [env.torch.cuda.Event(enable_timing=True) for _ in range(2)]
    stream = env.torch.cuda.ExternalStream(env.l.opf_stream())
    cold = []
    for _ in range(12):
        flush()
        with env.torch.cuda.stream(stream):
            ev[0].record()
        step()
        with env.torch.cuda.stream(stream):
            ev[1].record()
        ev[1].synchronize()
        cold.append(ev[0].elapsed_time(ev[1]))
    e2e_ms, nbytes = e2e_assign(env, u, expr, init, 20, 3)
    e2e_step = statistics.median(e2e_ms) / 20
    line = {"metric": METRIC, "value": updates / (ms * 1e-3) / 1e9, "unit": "GLUPS", "ms_per_step": ms,
            "config": {"workload": "FTCS2D heat equation 1025x1025 nodes FP64 5-point explicit (examples/FTCS2D/FTCS-OMP.cpp:26), Dirichlet 1",
                       "l2": "the 8.4 MB field is L2-resident in a real time loop, so it is timed that way; cold_ms_per_step = same launch after an L2 flush"},
            "cold_ms_per_step": statistics.median(cold), "step_kernel": kern,
            "graph_replay": {"ms_per_step": replay_ms, "value": updates / (replay_ms * 1e-3) / 1e9, "unit": "GLUPS",
                             "api": "opf_assign_repeat: the time loop as CUDA graphs of 32 steps (identical results, no launch gaps)"},
            "roofline": env.roofline(16.0, updates, kernel_ms, kname, note="L2-resident working set: the HBM roofline is not the binding limit here, launch latency is"),
            "e2e": {"value": updates / (e2e_step * 1e-3) / 1e9, "unit": "GLUPS", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "ms_per_step": e2e_step}}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_explicit("ftcs2d", 1025, 200, "ftcs2d 1025^2 nodes")
    return line


# ---- C3: CONV1D WENO5 on 2^26 cells; C3B: the batched variant of SURVEY 8d (8193 x 8192 field, dx<> only: 8192 independent lines)
def run_c3(env, args, compact=False, batched=False):
    import numpy as np
    from opflow_b200 import host
    from opflow_b200.host import D1WENO53Downwind, dx
    if batched:
        nx, ny = 8193, 8192
        mesh = host.MeshBuilder(2).newMesh(nx, ny + 1).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).build()
        b = host.ExprBuilder().setName("u").setMesh(mesh).setLoc([0, 1])
        b.setBC(0, 0, host.BCType.Dirc, 0.).setBC(0, 1, host.BCType.Dirc, 0.).setExt(0, 0, 3).setExt(0, 1, 3)
        u = b.build()
        lr = u.localRange
        x = np.linspace(0., 1., nx)
        line0 = np.where((x >= 0.2) & (x <= 0.4), 1.0, 0.0)
        init = np.asfortranarray(np.repeat(line0[:, None], lr.shape(2)[1], axis=1))
        u.from_numpy(init)
        c = 0.5 / (nx - 1)
        w = [min(u.assignableRange.end[d], lr.end[d]) - max(u.assignableRange.start[d], lr.start[d]) for d in range(2)]
        updates = w[0] * w[1]
        what = "CONV1D WENO5 advection, batched: 8193-node lines x 8192 lines FP64, u = u - dt*c*dx<D1WENO53Downwind>(u) (SURVEY 8d batched alternative)"
    else:
        n = 2 ** 26 + 1
        mesh = host.MeshBuilder(1).newMesh(n).setMeshOfDim(0, 0., 1.).build()
        u = host.ExprBuilder().setName("u").setMesh(mesh).setBC(0, 0, host.BCType.Dirc, 0.).setBC(0, 1, host.BCType.Dirc, 0.).setExt(3).build()
        x = np.linspace(0., 1., n)
        init = np.where((x >= 0.2) & (x <= 0.4), 1.0, 0.0)  # CONV1D.cpp:18
        u.from_numpy(init)
        c = 0.5 / (n - 1)
        updates = n - 2
        what = "CONV1D WENO5 advection, 2^26+1 nodes FP64, u = u - dt*c*dx<D1WENO53Downwind>(u), top-hat IC (examples/CONV1D/CONV1D.cpp:18-31)"
    expr = u - c * dx(D1WENO53Downwind, u)
    steps = 20 if compact else max(10, min(args.steps, 50))

    def step():
        u.assign(expr)
    for _ in range(5):
        step()
    t = env.timed(step, steps, 3 if compact else args.batches, barrier=False)
    ms = statistics.median(t) / steps
    kern = env.kernel_name()
    kernel_ms, kname = kernel_only(env, u, expr, steps)
    e2e_steps = 3
    e2e_ms, nbytes = e2e_assign(env, u, expr, init, e2e_steps, 2)
    e2e_step = statistics.median(e2e_ms) / e2e_steps
    line = {"metric": METRIC, "value": updates / (ms * 1e-3) / 1e9, "unit": "GLUPS", "ms_per_step": ms,
            "config": {"workload": what, "l2": "field (537 MB) is larger than L2 (126 MB): no flush needed"}, "step_kernel": kern,
            "roofline": env.roofline(16.0, updates, kernel_ms, kname,
                                     note="FP64-pipe bound, not HBM bound: ~100 DFMA-class instructions per cell after strength reduction (profiles/: "
                                          "sm__pipe_fp64_cycles_active of this kernel); the HBM fraction is reported as the contract asks"),
            "e2e": {"value": updates / (e2e_step * 1e-3) / 1e9, "unit": "GLUPS", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "ms_per_step": e2e_step}}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_explicit("weno_down", 2 ** 22 + 1, 20, "weno_down 2^22+1 nodes (bounded sample)")
    return line


# ---- C4: LidDriven2D pressure Poisson, 4096^2 cells
def run_c4(env, args, compact=False, n=4097):
    import numpy as np
    from opflow_b200 import host
    from opflow_b200.host import D2SecondOrderCentered as D2, EqnSolveHandler, StructSolverType as ST, d2x, d2y
    mesh = host.MeshBuilder(2).newMesh(n, n).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).build()

    def mk(name):
        b = host.ExprBuilder().setMesh(mesh).setName(name).setLoc([1, 1]).setExt(1)
        for d in range(2):
            b.setBC(d, 0, host.BCType.Neum, 0.).setBC(d, 1, host.BCType.Neum, 0.)
        return b.build()
    p, bf, pt = mk("p"), mk("b"), mk("pt")
    sh = pt.localRange.shape(2)
    xs = [(np.arange(sh[d]) + 0.5) / (n - 1) for d in range(2)]
    pt.from_numpy(np.asfortranarray(np.cos(2 * np.pi * xs[0])[:, None] * np.cos(np.pi * xs[1])[None, :]))

    def lap(f):
        return d2x(D2, f) + d2y(D2, f)
    bf.assign(lap(pt))
    relax = int(os.environ.get("OPF_C4_RELAX", "1"))  # PFMG relaxType of the reference's LidDriven2D.cpp:48 is 1 (Jacobi); 2 = symmetric red-black GS
    h = EqnSolveHandler(lambda e: (lap(e), bf), p, type_=ST.PCG, precond=ST.PFMG, tol=1e-10, maxIter=100, pinValue=True, staticMat=True, relaxType=relax)
    state = {}

    def solve():
        p.assign(0.0)
        state["st"] = h.solve()
    for _ in range(3):
        solve()
    solves = 5 if compact else max(5, min(args.steps, 20))
    t = env.timed(solve, solves, 3, barrier=False)
    ms = statistics.median(t) / solves
    st = state["st"]
    cells = (n - 1) ** 2
    # error against the manufactured solution (both shifted to p[first cell] = 0: the pinned cell, HYPREEqnSolveHandler.hpp:145-163)
    got, want = p.to_numpy(), pt.to_numpy()
    err = float(np.abs((got - got[0, 0]) - (want - want[0, 0])).max())
    # e2e: right-hand side from a pinned host buffer, solve, pressure back to the host
    hb, hp = host_pinned(env, sh), host_pinned(env, sh)
    hb.numpy()[...] = np.ascontiguousarray(bf.to_numpy().T)
    lrp = p.localRange

    def e2e_solve():
        bf.upload_raw(hb.data_ptr(), lrp)
        bf.updatePadding()
        solve()
        p.download_raw(hp.data_ptr(), lrp)
    e2e_solve()
    te = env.timed(e2e_solve, 3, 2, barrier=False)
    e2e = statistics.median(te) / 3
    gbs = 200.0 * cells * st.niter / (ms * 1e-3) / 1e9
    line = {"metric": "Poisson solve ms", "value": ms, "unit": "ms", "higher_is_better": False, "ms_per_step": ms, "iterations": st.niter, "relres": st.relerr,
            "levels": h.levels(), "max_abs_error_vs_manufactured": err, "cell_iterations_per_s": cells * st.niter / (ms * 1e-3),
            "config": {"workload": f"LidDriven2D pressure Poisson {n - 1}^2 cells (LidDriven2D.cpp:67-74): d2x(e)+d2y(e) == b, Neumann + pinValue, staticMat, tol 1e-10",
                       "solver": "matrix-free PCG + geometric multigrid V(1,1), weighted Jacobi", "l2": "vectors of 134 MB each exceed L2"},
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": env.peak, "unit": "GB/s", "frac": gbs / env.peak, "traffic": None, "kernel": "whole PCG iteration (operator, dots, "
                         "axpys, V-cycle)", "bytes_per_update": 200.0, "note": "aggregate of SURVEY 8d: ~200 B per fine cell per PCG iteration with a V(1,1) cycle"},
            "e2e": {"value": e2e, "unit": "ms", "h2d_bytes_per_step": hb.numel() * 8, "d2h_bytes_per_step": hp.numel() * 8}}
    if not args.no_cpu_baseline:
        cores = host_cores()
        try:
            t0 = time.time()
            r = ref_poisson(1025, 2, cores)
            if r:
                line["cpu_baseline"] = {"value": r["ms_per_solve"] * (float(n - 1) / 1024.0) ** 2, "unit": "ms", "cores": cores, "kind": "reference",
                                        "sample": f"reference pressure handler (HYPRE GMRES + PFMG) at 1024^2 cells: {r['ms_per_solve']:.0f} ms per solve ({r['niter']} outer "
                                                  f"iterations), scaled by the cell ratio to {n - 1}^2; {cores} threads, wall {time.time() - t0:.1f}s"}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "ms", "cores": cores, "kind": "reference", "sample": f"failed: {e}"[:160]}
    del h
    return line


def finish_line(line, env, args, launches=None):
    base = {"n_gpus": env.world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic"}
    for k, v in base.items():
        line.setdefault(k, v)
    line.setdefault("cpu_baseline", None)
    line["config"].setdefault("mode", args.mode)
    line["config"].setdefault("parallelism", "single GPU")
    if launches is not None:
        line.setdefault("gpu_launches", int(launches))
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batches", type=int, default=5)
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C3B", "C4", "C5"])
    ap.add_argument("--n", type=int, default=513, help="C2: nodes per axis (513 -> 511^3 updates per step)")
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the compact C1/C3/C4 block of the default line")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    args.batches = max(1, args.batches)
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and "NCCL_DEBUG" not in os.environ:
        # transport evidence for the line's comm_log_tail (which path NCCL chose for the halo messages); left alone if the caller set it
        os.environ["NCCL_DEBUG"] = "INFO"
        os.environ["NCCL_DEBUG_SUBSYS"] = "INIT,P2P,SHM,NET"
        os.environ["NCCL_DEBUG_FILE"] = os.environ["OPF_BENCH_NCCL_LOG"] = "/tmp/opf_bench_nccl_%p.log"
    if args.config == "C5":  # the C++ program owns the GPUs: no engine context in this process
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_c5
        line = bench_c5.run(args)
        if line is not None:
            print(json.dumps(line), flush=True)
        return 0
    from opflow_b200 import capi, host
    env = Env(args)
    host.set_mode(capi.MODE_FAST if args.mode == "fast" else capi.MODE_EXACT)
    l0 = env.l.opf_launch_count()
    if args.config == "C2":
        line = run_c2(env, args)
        if line is not None and env.world == 1 and not args.no_configs:
            blk = {}
            for name, fn in (("C1", run_c1), ("C3", run_c3), ("C4", run_c4)):
                try:
                    sub = fn(env, args, compact=True)
                    blk[name] = sub
                except Exception as e:  # a failing side config must not cost the headline line
                    blk[name] = {"error": f"{e}"[:200]}
            line["configs"] = blk
    else:
        if env.world > 1:
            raise SystemExit(f"--config {args.config} is a single-GPU configuration (multi-GPU: C2 and C5)")
        sampler = ClockSampler(env.local_rank)
        sampler.start()
        time.sleep(1.0)
        line = {"C1": run_c1, "C3": run_c3, "C3B": lambda e, a: run_c3(e, a, batched=True), "C4": run_c4}[args.config](env, args)
        line["clocks"] = sampler.stop()
        line = finish_line(line, env, args, env.l.opf_launch_count() - l0)
    if env.rank == 0 and line is not None:
        print(json.dumps(finish_line(line, env, args)), flush=True)
    if env.dist:
        capi.check(env.l.opf_comm_finalize())
        env.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
