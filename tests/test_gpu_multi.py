"""Runs tests/mgpu_check.py on 2 GPUs when the box has them (gpurun --gpus 2); skipped on single-GPU boxes."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_slab_decomposition_matches_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_frontend_program_on_two_gpus_matches_reference(tmp_path):
    """The FTCS-MPI.cpp set-up (cell-centred, ext 1, padding 1, EvenSplitStrategy) through the C++ front-end, one process per GPU:
    InitEnvironment builds the NCCL communicator, every rank dumps its block; the union equals the single-rank run of the
    unmodified reference (tests/golden/ftcs2d_mpi_n65_sin_s200.opfd) bit for bit in EXACT mode."""
    import json
    import numpy as np
    import torch
    from oracle import oracle as O
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    exe = os.path.join(ROOT, "tests", "frontend", "_bin", "fe_explicit")
    man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))["ftcs2d_mpi_n65_sin_s200"]
    args = [a for a in man["args"]]
    args[args.index("--ghosts") + 1] = "0"
    out = str(tmp_path / "block.opfd")
    # one process per GPU, the environment a launcher (torchrun / mpirun / srun) would provide; no MPI, no torch inside the program
    procs = []
    for rank in range(2):
        env = dict(os.environ, OPF_MODE="exact", OPF_RENDEZVOUS_DIR=str(tmp_path), RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT="29519")
        procs.append(subprocess.Popen([exe, *args, "--dump", out], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
    for pr in procs:
        so, se = pr.communicate(timeout=600)
        assert pr.returncode == 0, so[-2000:] + se[-2000:]
    s, e, ref = O.read_opfd(os.path.join(ROOT, "tests", "golden", "ftcs2d_mpi_n65_sin_s200.opfd"))
    covered = np.zeros(ref.shape, dtype=bool)
    for rank in range(2):
        bs, be, blk = O.read_opfd(out + f".{rank}")
        sl = tuple(slice(bs[d] - s[d], be[d] - s[d]) for d in range(2))
        assert np.array_equal(blk, ref[sl]), f"rank {rank}: max abs diff {np.abs(blk - ref[sl]).max()}"
        covered[sl] = True
    assert covered[1:-1, 1:-1].all()  # the two blocks tile the 64 x 64 cells
