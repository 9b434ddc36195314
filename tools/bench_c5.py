#!/usr/bin/env python
"""tools/bench_c5.py -- BASELINE config C5: 3-D Taylor-Green vortex, 1024 x 1024 x 128 cells per GPU in z-slabs (1024^3 on 8 GPUs).

The program that runs is tests/frontend/_bin/fe_tg3d: oracle/ref_drivers/ref_tg3d.cpp -- the operator set and time step of
examples/LidDriven/LidDriven3D.cpp:59-160 on the periodic set-up of examples/TaylorGreen/TGMPI.cpp:35-49 -- compiled by nvcc against
the B200 front-end (`#include <OpFlow>`), one process per GPU, InitEnvironment -> NCCL communicator, SlabSplitStrategy.  The same
source compiled against the unmodified reference is the parity oracle (tests/test_gpu_frontend.py::test_taylor_green_3d_matches_reference).
bench.py --config C5 calls run(): every rank (torchrun) starts the program with its own RANK / LOCAL_RANK; rank 0 relays the line."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "frontend", "_bin", "fe_tg3d")


def run(args):
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    n = int(os.environ.get("OPF_C5_N", "1025"))
    nz_cells = int(os.environ.get("OPF_C5_NZ", "128")) * world
    steps = max(1, min(args.steps, int(os.environ.get("OPF_C5_STEPS", "3"))))
    tol = os.environ.get("OPF_C5_TOL", "1e-8")
    if not os.path.exists(EXE):
        if rank == 0:
            print(json.dumps({"metric": "grid-point updates/sec (GLUPS)", "value": None, "unavailable": "tests/frontend/_bin/fe_tg3d not built (make -C tests/frontend tg3d)"}))
        return None
    env = dict(os.environ, OPF_MODE=args.mode, OPF_JOB_ID=f"benchc5-{os.environ.get('MASTER_PORT', '0')}-{os.getppid()}")
    t0 = time.time()
    r = subprocess.run([EXE, "--n", str(n), "--nz", str(nz_cells + 1), "--steps", str(steps), "--tol", tol], capture_output=True, text=True, env=env, timeout=3000)
    if r.returncode != 0:
        sys.stderr.write(r.stderr[-3000:])
        raise SystemExit(f"fe_tg3d failed on rank {rank} ({r.returncode})")
    if rank != 0:
        return None
    info = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    cells = info["cells"]
    exp_ms, mom_ms, poi_ms = info["explicit_ms_per_step"], info["momentum_ms_per_step"], info["poisson_ms_per_step"]
    step_ms = exp_ms + mom_ms + poi_ms
    sweeps = info["explicit_sweeps_per_step"]
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    # algorithmic bytes of the nine explicit sweeps of one step (16-40 B per cell depending on operand count):
    # dv, du corrections (4 and 6 operand fields), u/v/w += (3 x 24 B), projection (3 x 24 B), p += dp (24 B)
    bytes_per_cell = (8 * 5) + (8 * 7) + 3 * 24 + 3 * 24 + 24
    gbs = bytes_per_cell * cells / world / (exp_ms * 1e-3) / 1e9
    return {"metric": "grid-point updates/sec (GLUPS)", "value": sweeps * cells / (exp_ms * 1e-3) / 1e9, "unit": "GLUPS", "n_gpus": world, "steps": steps,
            "warmup": 1, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"TaylorGreen 3D {n - 1}x{n - 1}x{nz_cells} cells FP64, MAC staggering, periodic, LidDriven3D operator set (semi-implicit momentum "
                                   f"solves + explicit corrections + pinned Poisson solve + projection), tol {tol}", "mode": args.mode,
                       "parallelism": f"z-slabs x{world} (SlabSplitStrategy), NCCL halo exchange + allreduce, one process per GPU" if world > 1 else "single GPU",
                       "program": "oracle/ref_drivers/ref_tg3d.cpp through <OpFlow> of opflow_b200/include (tests/frontend/_bin/fe_tg3d)",
                       "l2": "every field (1.1 GB per GPU) exceeds L2"},
            "explicit_ms_per_step": exp_ms, "explicit_sweeps_per_step": sweeps, "momentum_ms_per_step": mom_ms, "momentum_iterations_per_step": info["momentum_iterations_per_step"],
            "poisson_ms_per_solve": poi_ms, "poisson_iterations": info["poisson_iterations_per_step"], "cell_steps_per_s": cells / (step_ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None, "kernel": "the nine explicit sweeps of a step (per GPU)",
                         "bytes_per_update": bytes_per_cell, "note": "host-clock timing around device-synchronised phases inside the program"},
            "cpu_baseline": None, "e2e": None, "gpu_launches": None, "wall_s": time.time() - t0}


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--mode", default="fast")
    line = run(ap.parse_args())
    if line:
        print(json.dumps(line))
