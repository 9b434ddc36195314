#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 evaluation engine for OpFlow's stencil hot path.

Workload (BASELINE.json configs[1]): 3-D FTCS heat equation, 513^3 nodes (511^3 interior updates per step), FP64,
7-point stencil  u = u + dt*alpha*(d2x(u) + d2y(u) + d2z(u))  with D2SecondOrderCentered, Dirichlet 1 on every face --
the 3-D extension of examples/FTCS2D/FTCS-OMP.cpp:26.  A "step" is one such assignment over the
whole field including the reference's implied updatePadding().

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n NODES] [--mode fast|exact] [--impl reference]

Prints ONE JSON line (rank 0).  value = grid-point updates/s (GLUPS) with the field resident in HBM; e2e = the same
metric through the C-ABI with host buffers (H2D of the input field and D2H of the result every step, pinned memory).
roofline: algorithmic bytes (16 B per update) / measured kernel time vs MEASURED_PEAKS.json hbm_gbs.
cpu_baseline: the reference's own CPU path (oracle/_ref, TBB, all host threads) on a bounded sample.
N > 1 (torchrun): weak scaling, z-slabs of 512 cells per GPU, NCCL halo exchange, no other collective.
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

BYTES_PER_UPDATE = 16.0  # read u (8) + write u' (8): SURVEY.md section 8d


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
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ref_sample(n, steps, warmup, threads):
    """the reference's own CPU implementation (oracle/_ref/bin/ref_explicit, unmodified OpFlow + TBB)"""
    from oracle import oracle as O
    if not O.ref_available("ref_explicit"):
        return None
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(O.HERE, "_ref", "lib") + ":" + env.get("LD_LIBRARY_PATH", "")
    r = O.run_ref("ref_explicit", "--case", "ftcs3d", "--n", n, "--steps", steps, "--warmup", warmup, "--threads", threads,
                  "--init", "sin", env=env, timeout=1500)[0]
    return r


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: times the reference CPU path on this box's host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    n = args.n
    # bounded sample: the reference sustains ~0.19 GLUPS on 16 cores, so K + W steps of the full 513^3 problem take (K + W) x 0.7 s;
    # above ~2 minutes the sample per step shrinks (same expression, same arithmetic, smaller cube) -- the metric is a rate
    budget_cells = 120.0 * 0.012e9 * cores / max(1, args.steps + args.warmup)
    if (n - 2) ** 3 > budget_cells:
        n = max(129, min(n, int(round(budget_cells ** (1.0 / 3.0))) + 2))
    r = None
    t0 = time.time()
    try:
        r = ref_sample(n, args.steps, args.warmup, cores)
    except Exception as e:  # e.g. not enough host memory for 3 x 2 GiB: fall back to a smaller sample
        sys.stderr.write(f"reference at n={n} failed ({e}); retrying at n=257\n")
    if r is None:
        try:
            n = 257
            r = ref_sample(n, args.steps, args.warmup, cores)
        except Exception as e:
            print(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref/bin/ref_explicit not runnable: {e}"[:200]}))
            return 0
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/bin/ref_explicit missing (run oracle/build_ref.sh)"}))
        return 0
    glups = r["mlups"] / 1e3
    sample = (f"ftcs3d {n}^3 nodes per step" + ("" if n == args.n else f" (bounded sample of the {args.n}^3 workload)")
              + f", {args.steps} steps after {args.warmup} warm-up, {cores} TBB threads, wall {time.time() - t0:.1f}s")
    line = {"impl": "reference", "metric": "grid-point updates/sec (GLUPS)", "value": glups, "unit": "GLUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds"] / max(1, args.steps) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"FTCS3D heat equation {args.n}x{args.n}x{args.n} nodes FP64 7-point explicit, u = u + dt*alpha*(d2x+d2y+d2z)(u), Dirichlet 1",
                       "mode": "reference CPU path (unmodified OpFlow, TBB rangeFor, all host cores)", "parallelism": "host cores only"},
            "cpu_baseline": {"value": glups, "unit": "GLUPS", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": glups, "unit": "GLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--n", type=int, default=513, help="nodes per axis (513 -> 511^3 updates per step)")
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch

    from opflow_b200 import capi, host
    from opflow_b200.host import D2SecondOrderCentered as D2, d2x, d2y, d2z

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: opflow_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    l = capi.lib()
    capi.check(l.opf_init(local_rank))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = (C.c_ubyte * 128)()
            capi.check(l.opf_comm_unique_id(raw))
            idbuf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        raw = (C.c_ubyte * 128)(*idbuf.cpu().tolist())
        capi.check(l.opf_comm_init(rank, world, raw))
    host.set_mode(capi.MODE_FAST if args.mode == "fast" else capi.MODE_EXACT)

    n = args.n
    nz_cells = (n - 1) * world  # weak scaling: 512 z-cells per GPU
    mesh = host.MeshBuilder(3).newMesh(n, n, nz_cells + 1).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.) \
        .setMeshOfDim(2, 0., float(world)).build()
    b = host.ExprBuilder().setName("u").setMesh(mesh)
    for d in range(3):
        b.setBC(d, host.DimPos.start, host.BCType.Dirc, 1.).setBC(d, host.DimPos.end, host.BCType.Dirc, 1.)
    if world > 1:
        b.setPadding(1).setSplitStrategy(world, rank, host.split_slab(mesh, world))
    u = b.build()
    u.assign(0.0)
    # non-trivial initial condition (deterministic; same on every run): product of sines, evaluated with numpy on host
    lr = u.localRange
    xs = [np.linspace(0.0, 1.0 if d < 2 else float(world), (n if d < 2 else nz_cells + 1))[lr.start[d]:lr.end[d]] for d in range(3)]
    init = (np.sin(np.pi * xs[0])[:, None, None] * np.sin(np.pi * xs[1])[None, :, None] * np.sin(np.pi * xs[2])[None, None, :])
    u.from_numpy(np.asfortranarray(init))
    dt = 0.1 / (n - 1) ** 2
    c = dt * 1.0
    expr = u + c * (d2x(D2, u) + d2y(D2, u) + d2z(D2, u))
    w = [min(u.assignableRange.end[d], lr.end[d]) - max(u.assignableRange.start[d], lr.start[d]) for d in range(3)]
    updates_per_step_rank = w[0] * w[1] * w[2]

    def barrier():
        capi.check(l.opf_synchronize())
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        u.assign(expr)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = l.opf_launch_count()
    ms = C.c_float()
    capi.check(l.opf_timer_begin())
    for _ in range(args.steps):
        u.assign(expr)
    capi.check(l.opf_timer_end(C.byref(ms)))
    barrier()
    launches = l.opf_launch_count() - launches0
    t_ms = float(ms.value)

    # dominant kernel alone: the same aliased assignment without the trailing updatePadding() (OPF_ASSIGN_NO_PADDING), i.e. exactly
    # one tma_kernel launch per iteration (ping-pong buffers as in a real step), timed with CUDA events on the engine's stream
    sig, fields, scalars = expr.flatten()
    F = (C.c_void_p * len(fields))(*[f.h for f in fields])
    S = (C.c_double * len(scalars))(*scalars)
    NO_PADDING = 1
    for _ in range(3):
        capi.check(l.opf_assign_ex(u.h, capi.OP_EQ, sig.encode(), F, len(fields), S, len(scalars), NO_PADDING))
    capi.check(l.opf_synchronize())
    capi.check(l.opf_timer_begin())
    for _ in range(args.steps):
        capi.check(l.opf_assign_ex(u.h, capi.OP_EQ, sig.encode(), F, len(fields), S, len(scalars), NO_PADDING))
    capi.check(l.opf_timer_end(C.byref(ms)))
    kernel_ms = float(ms.value) / args.steps
    u.updatePadding()

    # ---- e2e: host buffers through the C ABI (pinned), H2D input + step + D2H result inside the timed region
    e2e_steps = max(1, args.e2e_steps)
    shape = lr.shape(3)
    hin = torch.empty(shape[::-1], dtype=torch.float64).pin_memory()
    hout = torch.empty(shape[::-1], dtype=torch.float64).pin_memory()
    hin.numpy()[...] = np.ascontiguousarray(init.transpose(2, 1, 0))
    nbytes = hin.numel() * 8
    # the public host-buffer call: opf_assign_host pipelines upload | sweep | download in z-slabs (PCIe both ways at once)
    def e2e_step():
        capi.check(l.opf_assign_host(u.h, capi.OP_EQ, sig.encode(), F, len(fields), S, len(scalars), u.h, C.c_void_p(hin.data_ptr()),
                                     C.c_void_p(hout.data_ptr())))

    for _ in range(2):
        e2e_step()
    barrier()
    capi.check(l.opf_timer_begin())
    for _ in range(e2e_steps):
        e2e_step()
    capi.check(l.opf_timer_end(C.byref(ms)))
    barrier()
    e2e_ms = float(ms.value)
    clocks = sampler.stop() if rank == 0 else None  # sampled across the timed steps, the kernel-only loop and the e2e leg

    if world > 1:
        t = torch.tensor([t_ms, e2e_ms, kernel_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms, e2e_ms, kernel_ms = t.tolist()
        tot = torch.tensor([float(updates_per_step_rank), float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        updates_per_step = tot[0].item()
        launches_total = int(tot[1].item())
    else:
        updates_per_step = float(updates_per_step_rank)
        launches_total = int(launches)

    if rank == 0:
        ms_per_step = t_ms / args.steps
        glups = updates_per_step / (ms_per_step * 1e-3) / 1e9
        e2e_glups = updates_per_step / (e2e_ms / e2e_steps * 1e-3) / 1e9
        peak, peak_src = peaks()
        # roofline of the dominant kernel: algorithmic bytes of ONE launch (16 B x this rank's updates) / its measured duration
        achieved = BYTES_PER_UPDATE * updates_per_step_rank / (kernel_ms * 1e-3) / 1e9
        step_achieved = BYTES_PER_UPDATE * updates_per_step_rank / (ms_per_step * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": "opf::tma_kernel<Add<F0,Mul<S0,Add<Add<D2C<0,F1>,D2C<1,F2>>,D2C<2,F3>>>>, Fast, alias0, CX=2, 64x4 threads, ring 6, uniform>",
                "kernel_ms": kernel_ms, "peak_source": peak_src,
                "note": "achieved = 16 B x updates of one launch / mean launch duration (CUDA events on the engine stream, launches "
                        "back to back, no BC/halo launches between them); whole_step_* adds the BC-face fill (and halo) launches",
                "whole_step_achieved": step_achieved, "whole_step_frac": step_achieved / peak}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "ftcs3d_traffic.json")))
            roof["traffic"] = prof.get("dram_bytes_per_launch")
        except Exception:
            pass
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = host_cores()
            try:
                t0 = time.time()
                r = ref_sample(257, 100, 2, cores)
                if r:
                    cpu = {"value": r["mlups"] / 1e3, "unit": "GLUPS", "cores": cores, "kind": "reference",
                           "sample": f"reference ftcs3d 257^3 nodes, 100 steps after 2 warm-up, {cores} TBB threads, wall {time.time() - t0:.1f}s"}
            except Exception as e:
                cpu = {"value": None, "unit": "GLUPS", "cores": cores, "kind": "reference", "sample": f"failed: {e}"[:160]}
        line = {"metric": "grid-point updates/sec (GLUPS)", "value": glups, "unit": "GLUPS", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"FTCS3D heat equation {n}x{n}x{nz_cells + 1} nodes FP64 7-point explicit, "
                                       f"u = u + dt*alpha*(d2x+d2y+d2z)(u), Dirichlet 1",
                           "mode": args.mode, "l2": "field (1.1 GB per GPU) is larger than L2 (126 MB): no flush needed",
                           "parallelism": f"z-slabs x{world}" if world > 1 else "single GPU"},
                "roofline": roof, "cpu_baseline": cpu,
                "e2e": {"value": e2e_glups, "unit": "GLUPS", "h2d_bytes_per_step": nbytes * world, "d2h_bytes_per_step": nbytes * world,
                        "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps},
                "gpu_launches": launches_total, "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        capi.check(l.opf_comm_finalize())
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
