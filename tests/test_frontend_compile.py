"""CPU-side checks of the C++ front-end (opflow_b200/include/OpFlow): the reference's example programs parse and instantiate against
it unchanged (host-only g++ pass, no kernels), and a front-end program run on a box without a GPU fails loudly instead of falling
back to the CPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("OPF_REFERENCE", "/root/reference")
INC = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "opflow_b200", "include")]
EXAMPLES = ["FTCS2D/FTCS-OMP.cpp", "FTCS2D/FTCS-MPI.cpp", "FTCS/FTCS.cpp", "CONV1D/CONV1D.cpp", "LidDriven/LidDriven2D.cpp", "LidDriven/LidDriven2D-MPI.cpp",
            "LidDriven/LidDriven3D.cpp"]


@pytest.mark.parametrize("example", EXAMPLES)
def test_reference_example_compiles_unchanged(example):
    src = os.path.join(REF, "examples", example)
    if not os.path.exists(src):
        pytest.skip("reference tree not present (GPU box)")
    r = subprocess.run(["g++", "-std=c++20", "-fsyntax-only", "-x", "c++", *INC, src], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]


PROG = r"""
#include <OpFlow>
using namespace OpFlow;
int main() {
    using Mesh = CartesianMesh<Meta::int_<2>>;
    using Field = CartesianField<Real, Mesh>;
    auto mesh = MeshBuilder<Mesh>().newMesh(17, 17).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).build();   // host only: fine
    auto u = ExprBuilder<Field>().setName("u").setMesh(mesh).setBC(0, DimPos::start, BCType::Dirc, 1.).build();  // needs the device
    u = u + 0.1 * (d2x<D2SecondOrderCentered>(u) + d2y<D2SecondOrderCentered>(u));
    std::puts("computed without a GPU?!");
    return 0;
}
"""


def test_frontend_program_fails_loudly_without_a_gpu(tmp_path):
    """host-only build (g++) of a front-end program, linked against libopflow_b200.so, run with no visible device"""
    lib = os.path.join(ROOT, "opflow_b200", "libopflow_b200.so")
    if not os.path.exists(lib) or shutil.which("g++") is None:
        pytest.skip("library not built")
    src, exe = tmp_path / "p.cpp", tmp_path / "p"
    src.write_text(PROG)
    r = subprocess.run(["g++", "-std=c++20", "-O1", *INC, str(src), "-o", str(exe), "-L" + os.path.dirname(lib), "-lopflow_b200",
                        "-Wl,-rpath," + os.path.dirname(lib)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode != 0 and "computed without a GPU" not in r.stdout
    assert "no CUDA device" in r.stderr or "no CPU fallback" in r.stderr, r.stderr
