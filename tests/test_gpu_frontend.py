"""The C++ front-end (`#include <OpFlow>` of opflow_b200/include) as a drop-in for the reference's headers.

tests/frontend/Makefile compiles, with nvcc against the B200 front-end,
  * the SAME driver sources that oracle/build_ref.sh compiles against the unmodified reference (oracle/ref_drivers/*.cpp) --
    so their outputs must reproduce tests/golden/ (written by the reference build of those sources), and
  * the reference's own example programs, unchanged (examples/FTCS2D/FTCS-OMP.cpp, examples/CONV1D/CONV1D.cpp).
Bit-exact in OPF_MODE=exact (index maps, boundary classification, ghost values, explicit fields); FAST within 1e-12; implicit
solutions within 1e-10 relative (BASELINE.json north_star tolerances)."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "frontend", "_bin")
GOLD = os.path.join(ROOT, "tests", "golden")

pytestmark = pytest.mark.gpu


def run(exe, *args, mode="exact", cwd=None, timeout=900):
    path = os.path.join(BIN, exe)
    if not os.path.exists(path):
        pytest.fail(f"{path} missing: run `make -C tests/frontend` (done by __graft_entry__.build())")
    env = dict(os.environ, OPF_MODE=mode)
    r = subprocess.run([path, *map(str, args)], capture_output=True, text=True, timeout=timeout, env=env, cwd=cwd)
    assert r.returncode == 0, f"{exe} failed ({r.returncode}): {r.stderr[-2000:]}"
    return r


def same(a, b, path=""):
    """deep equality of two parsed JSON documents, floats compared exactly"""
    if isinstance(a, dict):
        assert isinstance(b, dict) and a.keys() == b.keys(), path
        for k in a:
            same(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, list):
        assert isinstance(b, list) and len(a) == len(b), path
        if a and all(isinstance(x, (int, float)) for x in a):
            assert np.array_equal(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)), \
                f"{path}: max abs diff {np.abs(np.asarray(a, float) - np.asarray(b, float)).max()}"
        else:
            for i, (x, y) in enumerate(zip(a, b)):
                same(x, y, f"{path}[{i}]")
    else:
        assert a == b, f"{path}: {a!r} != {b!r}"


def test_fields_driver_reproduces_reference_json():
    """range classification table, ghost values of every BC kind (1-D and 2-D incl. corners), prepared-expression ranges / loc,
    EvenSplitStrategy maps: the whole document the reference build of ref_fields.cpp printed"""
    got = json.loads(run("fe_fields").stdout)
    ref = json.load(open(os.path.join(GOLD, "ref_fields.json")))
    # a side without a BC leaves its ghost cells unwritten: the reference prints whatever malloc returned there (denormal garbage
    # in the fixture), so on such sides only the cells of localRange are comparable (same rule as tests/test_gpu_golden.py)
    for g, r in zip(got["ranges1d"], ref["ranges1d"]):
        fr, loc = r["field"]["range"], r["ranges"]["local"]
        lo = loc[0][0] if r["bc"][0] == "Undefined" else fr[0][0]
        hi = loc[1][0] if r["bc"][1] == "Undefined" else fr[1][0]
        for doc in (g, r):
            doc["field"]["values"] = doc["field"]["values"][lo - fr[0][0]:hi - fr[0][0]]
    same(got, ref)


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_explicit_driver_reproduces_reference_dumps(mode, tmp_path):
    man = json.load(open(os.path.join(GOLD, "manifest.json")))
    for name, info in man.items():
        out = tmp_path / (name + ".opfd")
        r = run("fe_explicit", *info["args"], "--dump", out, mode=mode)
        line = json.loads(r.stdout.strip().splitlines()[-1])
        for k, v in info["stdout"].items():
            if k != "threads":
                assert line[k] == v, (name, k)
        s, e, ref = O.read_opfd(os.path.join(GOLD, name + ".opfd"))
        s2, e2, got = O.read_opfd(str(out))
        assert (s, e) == (s2, e2), name
        if mode == "exact":
            assert np.array_equal(got, ref), f"{name}: max abs diff {np.abs(got - ref).max()}"
        else:
            tol = 1e-11 if "weno" in name else 1e-12
            assert np.abs(got - ref).max() <= tol * max(np.abs(ref).max(), 1e-300), name


def test_implicit_driver_matches_reference_solves():
    """Solve(L_h(e) == b) through the front-end's EqnSolveHandler vs HYPRE GMRES+PFMG solves of the reference: same right-hand
    side bit for bit (explicit path), solution within 1e-10 relative L-inf"""
    got = json.loads(run("fe_implicit").stdout)["cases"]
    ref = json.load(open(os.path.join(GOLD, "ref_implicit.json")))["cases"]
    assert [c["name"] for c in got] == [c["name"] for c in ref]
    for g, r in zip(got, ref):
        assert g["range"] == r["range"], g["name"]
        assert np.array_equal(np.asarray(g["b"]), np.asarray(r["b"])), g["name"] + ": right-hand side differs"
        gp, rp = np.asarray(g["p"]), np.asarray(r["p"])
        err = np.abs(gp - rp).max() / max(np.abs(rp).max(), 1e-300)
        assert err <= 1e-10, f"{g['name']}: solution differs by {err:.3e} relative"
        assert g["relerr"] <= 1e-12, g["name"]


def test_reference_example_ftcs_omp_unchanged(tmp_path):
    """examples/FTCS2D/FTCS-OMP.cpp compiled unchanged: 1025^2, 5000 steps; its printed centre value equals the reference's"""
    exe = os.path.join(BIN, "ref_FTCS-OMP")
    if not os.path.exists(exe):
        pytest.skip("reference tree was not available when the front-end programs were built")
    r = run("ref_FTCS-OMP", cwd=tmp_path)
    m = re.search(r"Center val: ([-+0-9.eE]+)", r.stderr + r.stdout)
    assert m, r.stderr[-500:]
    gold = json.load(open(os.path.join(GOLD, "examples.json")))["ftcs_omp_1025_5000_center"]
    assert float(m.group(1)) == gold, (m.group(1), gold)


def test_reference_example_conv1d_unchanged(tmp_path):
    """examples/CONV1D/CONV1D.cpp compiled unchanged: its last Tecplot zone equals the upwind1_n101_s100 fixture (6 digits)"""
    exe = os.path.join(BIN, "ref_CONV1D")
    if not os.path.exists(exe):
        pytest.skip("reference tree was not available when the front-end programs were built")
    run("ref_CONV1D", cwd=tmp_path)
    vals = last_zone(tmp_path / "u.tec")
    s, e, ref = O.read_opfd(os.path.join(GOLD, "upwind1_n101_s100.opfd"))
    core = ref[(0 - s[0]):(101 - s[0])]  # the example writes after every step: the last zone is the state after 100 steps
    assert vals.shape == core.shape
    assert np.abs(vals - core).max() <= 1e-9 * max(np.abs(core).max(), 1.0)


def last_zone(path):
    """values of the last zone of a reference-layout Tecplot file (ORDERED / BLOCK: dim coordinate blocks, then the variable)"""
    txt = open(path).read()
    dim = txt.split("VARIABLES = ", 1)[1].split("\n", 1)[0].count('"') // 2 - 1
    lines = txt.rsplit("ZONE\n", 1)[1].splitlines()
    ext = [int(x) for x in re.findall(r"= (\d+)", lines[1])]
    n = int(np.prod(ext))
    body = lines[3:]
    shared = body[0].startswith("VARSHARELIST")
    data = np.array([float(x) for x in (body[1:1 + n] if shared else body[dim * n:(dim + 1) * n])])
    return data.reshape(ext, order="F")


def test_reference_example_liddriven2d_unchanged(tmp_path):
    """examples/LidDriven/LidDriven2D.cpp compiled unchanged (projection method: two non-symmetric implicit momentum solves with
    field-dependent coefficients + the pinned Neumann pressure Poisson solve per step, 65^2, 1000 steps): final u, v, p against
    the reference's own run (HYPRE GMRES+PFMG at 1e-10).  The flow reaches its steady state, so the comparison is at the level
    of the solver tolerance accumulated over the run, not of round-off."""
    exe = os.path.join(BIN, "ref_LidDriven2D")
    if not os.path.exists(exe):
        pytest.skip("ref_LidDriven2D not built (make -C tests/frontend liddriven)")
    run("ref_LidDriven2D", cwd=tmp_path, mode="fast", timeout=3000)
    gold = json.load(open(os.path.join(GOLD, "liddriven2d_n65_s1000.json")))
    for name in "uvp":
        g = gold[name]
        ref = np.asarray(g["values"]).reshape([g["I"], g["J"]], order="F")
        got = last_zone(tmp_path / f"{name}.tec")
        assert got.shape == ref.shape, name
        if name == "p":  # defined up to the pinned constant; both pin the same cell, still compare mean-free
            got, ref = got - got.mean(), ref - ref.mean()
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err <= 1e-6, f"{name}: relative L-inf difference {err:.3e}"


def test_host_access_semantics_program():
    """tests/frontend/fe_hostaccess.cpp (host lambdas reading/writing device fields through operator[], copies, compound ops,
    conditional, initBy): the same source prints PASS when built against the unmodified reference on the CPU (checked in the
    authoring container with the oracle/build_ref.sh flags) -- here built against the B200 front-end"""
    r = run("fe_hostaccess", mode="exact")
    assert r.stdout.strip().endswith("PASS"), r.stdout[-2000:]


def test_ops_driver_reproduces_reference_json():
    """oracle/ref_drivers/ref_ops.cpp (flux-limiter interpolators of all ten schemes in both directions, 3x3 and 5x3 convolutions)
    compiled against the B200 front-end: in EXACT mode its output -- ranges, locations and every value as a hex double -- is the
    byte-for-byte document the unmodified reference printed (tests/golden/ref_ops.json); FAST within 1e-12"""
    ref_txt = open(os.path.join(GOLD, "ref_ops.json")).read()
    assert run("fe_ops", mode="exact").stdout == ref_txt
    got, ref = json.loads(run("fe_ops", mode="fast").stdout), json.loads(ref_txt)
    for g, r in zip(got["cases"], ref["cases"]):
        assert g["acc"] == r["acc"] and g["loc"] == r["loc"]
        a, b = (np.array([float.fromhex(v) for v in c["val"]]) for c in (g, r))
        assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max(), g["node"]


def test_poisson_driver_matches_reference_solution(tmp_path):
    """oracle/ref_drivers/ref_poisson.cpp (config C4's pressure handler: Neumann, pinValue, staticMat, GMRES + PFMG request) at 256^2
    cells against the solution the unmodified reference computed with HYPRE (both driven to 1e-13): <= 1e-10 relative"""
    out = str(tmp_path / "p.opfd")
    r = run("fe_poisson", "--n", 257, "--solves", 2, "--tol", "1e-13", "--dump", out, mode="fast")
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["relerr"] <= 1e-12, info
    _, _, got = O.read_opfd(out)
    _, _, ref = O.read_opfd(os.path.join(GOLD, "poisson_n257.opfd"))
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err <= 1e-10, f"pressure differs by {err:.3e} relative"


def test_reference_example_ftcs_mpi_unchanged(tmp_path):
    """examples/FTCS2D/FTCS-MPI.cpp compiled unchanged, run as a single worker (its set-up on 2 GPUs is covered bit-exactly by
    tests/test_gpu_multi.py through the shared driver source): it must run to completion and write its output"""
    exe = os.path.join(BIN, "ref_FTCS-MPI")
    if not os.path.exists(exe):
        pytest.skip("reference tree was not available when the front-end programs were built")
    r = run("ref_FTCS-MPI", cwd=tmp_path, timeout=1800)
    assert r.returncode == 0


@pytest.mark.parametrize("mode,tol_uvw,tol_p", [("exact", 1e-9, 1e-8), ("fast", 1e-9, 1e-8)])
def test_taylor_green_3d_matches_reference(tmp_path, mode, tol_uvw, tol_p):
    """BASELINE config C5's program (oracle/ref_drivers/ref_tg3d.cpp: LidDriven3D.cpp's operator set and time step on the periodic
    Taylor-Green box) at 32^3 cells, 1 + 2 steps, against the run of the unmodified reference (HYPRE GMRES / PCG + PFMG), all solves
    driven to 1e-12: velocities and pressure agree to the accumulated solver tolerance (three semi-implicit momentum solves, the
    explicit corrections, the pinned periodic Poisson solve and the projection per step)"""
    exe = os.path.join(BIN, "fe_tg3d")
    if not os.path.exists(exe):
        pytest.skip("fe_tg3d not built (make -C tests/frontend tg3d: ~30 minutes of nvcc)")
    pre = str(tmp_path / "tg")
    r = run("fe_tg3d", "--n", 33, "--steps", 2, "--tol", "1e-12", "--dump", pre, mode=mode, timeout=1800)
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["cells"] == 32 ** 3
    for name, tol in (("u", tol_uvw), ("v", tol_uvw), ("w", tol_uvw), ("p", tol_p)):
        _, _, got = O.read_opfd(pre + f"_{name}.opfd")
        _, _, ref = O.read_opfd(os.path.join(GOLD, f"tg3d_n33_s2_{name}.opfd"))
        scale = max(np.abs(ref).max(), 1e-3)
        if name == "p":
            got, ref = got - got.mean(), ref - ref.mean()
        err = np.abs(got - ref).max() / scale
        assert err <= tol, f"{name}: relative L-inf difference {err:.3e} > {tol}"


def test_binary_streams_round_trip(tmp_path):
    """tests/frontend/fe_io.cpp: RawBinaryOStream (.cart files in the reference's layout) and H5Stream written from asynchronous device
    snapshots inside a time loop, read back by RawBinaryIStream / H5Stream(StreamIn): bit-identical; and the .cart header is the
    reference's (RawBinaryStream.hpp:98-160: name, dim, nproc, time, mesh range, coordinates, accessible and local range)"""
    import struct
    os.makedirs(tmp_path / "out")
    r = run("fe_io", cwd=tmp_path)
    assert r.stdout.strip().endswith("PASS"), r.stdout[-500:] + r.stderr[-500:]
    b = open(tmp_path / "out" / "u_2.cart", "rb").read()
    n = struct.unpack_from("i", b, 0)[0]
    assert b[4:4 + n] == b"u"
    dim, nproc = struct.unpack_from("ii", b, 4 + n)
    t = struct.unpack_from("d", b, 12 + n)[0]
    assert (dim, nproc, t) == (2, 1, 2.0)
    assert struct.unpack_from("4i", b, 20 + n) == (0, 33, 0, 17)
    off = 20 + n + 16 + 8 * (33 + 17)
    assert struct.unpack_from("8i", b, off) == (0, 33, 0, 17, 0, 33, 0, 17)
    assert len(b) == off + 32 + 8 * 33 * 17


def test_functor_adaptor_convolution_program(tmp_path):
    """oracle/ref_drivers/ref_adapt.cpp (the volume-constraint step of UniLS.cpp:104-118: UniOpAdaptor of smoothDelta, pow, two 3 x 3
    convolutions in one assignment) against the dump of the unmodified reference.  The functor calls cos and the expression calls
    pow: CUDA's and glibc's libm agree to 1-2 ulp, not bit for bit, hence 1e-13 relative in both modes."""
    out = str(tmp_path / "p.opfd")
    _, _, ref = O.read_opfd(os.path.join(GOLD, "adapt_n65.opfd"))
    for mode in ("exact", "fast"):
        run("fe_adapt", "--dump", out, mode=mode)
        _, _, got = O.read_opfd(out)
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err <= 1e-13, f"{mode}: relative L-inf difference {err:.3e}"


@pytest.mark.parametrize("n,stride", [(129, 2), (257, 4)])
def test_lid_driven_2d_at_multigrid_sizes_matches_reference(tmp_path, n, stride):
    """oracle/ref_drivers/ref_ld2d.cpp (examples/LidDriven/LidDriven2D.cpp:10-96 with n, steps, tol on the command line) at 128^2 and
    256^2 cells, 1 + 3 steps, against the run of the unmodified reference (HYPRE GMRES + PFMG on every equation).  At these sizes
    nu dt / h^2 is 0.8 / 3.3: the momentum operators  e/dt + conv(u, e) - nu/2 lap(e)  are no longer diagonally dominant and the
    engine preconditions them with the coefficient-carrying V-cycle (restricted u, v, du on every level) instead of Jacobi.
    Both sides iterate to a relative residual of 1e-10; u, v (O(1e-1)) and the mean-free pressure agree to 1e-8 of their scale."""
    exe = os.path.join(BIN, "fe_ld2d")
    if not os.path.exists(exe):
        pytest.skip("fe_ld2d not built (make -C tests/frontend liddriven: ~4 minutes of nvcc)")
    pre = str(tmp_path / "ld")
    r = run("fe_ld2d", "--n", n, "--steps", 3, "--tol", "1e-10", "--stride", stride, "--dump", pre, mode="fast")
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["cells"] == (n - 1) ** 2
    for name in "uvp":
        _, _, got = O.read_opfd(pre + f"_{name}.opfd")
        _, _, ref = O.read_opfd(os.path.join(GOLD, f"ld2d_n{n}_s3_{name}.opfd"))
        assert got.shape == ref.shape
        if name == "p":
            got, ref = got - got.mean(), ref - ref.mean()
        scale = max(np.abs(ref).max(), 1e-3)
        err = np.abs(got - ref).max() / scale
        assert err <= 1e-8, f"{name}: relative L-inf difference {err:.3e}"
    # the preconditioner did its work: Jacobi-preconditioned GMRES needs ~50 iterations per momentum solve at 256^2 (OPF_MG_COEF=0:
    # 99 per step), the V-cycle ~10.  At 128^2 the diagonal still dominates (row sum / diagonal = 0.38 >= 1/4) and the engine keeps Jacobi.
    if n >= 257:
        assert info["momentum_iterations_per_step"] <= 30, info
