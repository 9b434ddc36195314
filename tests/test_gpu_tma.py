"""The TMA tile skeleton (opf::tma_kernel -- the kernel bench.py times) under the oracle and the reference itself.

tma_kernel is selected for 3-D fields whose assigned x-extent is >= 64 (opf_device.cuh launch_assign); every case here asserts
through opf_last_kernel_name() that it is the kernel that actually ran, then compares with the oracle on the same seeded input:
EXACT mode bit-identical, FAST mode within 1e-12 relative L-inf (BASELINE.json north_star).  Reference behaviour:
FieldAssigner::assign_impl (src/Core/Loops/FieldAssigner.hpp:40-86)."""
import math
import os
import tempfile

import numpy as np
import pytest

from opflow_b200 import capi, host
from opflow_b200.host import D1FirstOrderCentered as D1, D2SecondOrderCentered as D2, d2x, d2y, d2z, dz
from helpers import assert_same, dirc, gpu_storage, make_pair, make_pair_on, set_both
from oracle import oracle as O

pytestmark = pytest.mark.gpu

MODES = [(capi.MODE_EXACT, True), (capi.MODE_FAST, False)]


def stretched(n, a=0.0, b=1.0):
    s = np.linspace(0, 1, n)
    return a + (b - a) * (s + 0.15 * np.sin(2 * np.pi * s) / (2 * np.pi))


def lap(f):
    return d2x(D2, f) + d2y(D2, f) + d2z(D2, f)


def kernel_name():
    return capi.lib().opf_last_kernel_name().decode()


def expect_tma():
    assert kernel_name() == "opf::tma_kernel", f"dispatch chose {kernel_name()}, the test does not cover the TMA skeleton"


@pytest.mark.parametrize("mode,exact", MODES)
@pytest.mark.parametrize("dims", [(97, 41, 29), (130, 19, 23), (257, 9, 40)])
def test_ftcs3d_uniform_ragged(engine, oracle, mode, exact, dims):
    """aliased 7-point FTCS on uniform meshes (UNI variant: constant-bank coefficients), ragged extents: partial x tiles, partial
    y tiles, z chunks shorter and longer than the 16-plane march"""
    host.set_mode(mode)
    g, o = make_pair(list(dims), [0, 0, 0], [1, 1.5, 0.75], bc=dirc(3))
    set_both(g, o)
    c = 2e-5
    e = g + c * lap(g)
    for _ in range(4):
        g.assign(e)
        expect_tma()
        oracle.assign(o, e.signature(), [o] * 4, [c])
    a, b = gpu_storage(g, o)
    assert_same(a, b, exact, what=f"tma ftcs3d {dims}")


@pytest.mark.parametrize("mode,exact", MODES)
@pytest.mark.parametrize("loc", [[0, 0, 0], [1, 1, 1]])
def test_ftcs3d_stretched_mixed_bc(engine, oracle, mode, exact, loc):
    """stretched mesh (per-axis coefficient arrays: the non-UNI variant), Corner and Center, Dirichlet / Neumann / Symm / ASymm, ext 1"""
    host.set_mode(mode)
    dims = [101, 37, 21]
    coords = [stretched(n) for n in dims]
    bc = {(0, 0): (capi.BC_DIRC, 1.0), (0, 1): (capi.BC_NEUM, 0.5), (1, 0): (capi.BC_NEUM, -0.25), (1, 1): (capi.BC_DIRC, 2.0),
          (2, 0): (capi.BC_SYMM, 0.0), (2, 1): (capi.BC_ASYMM, 0.0)}
    g, o = make_pair(dims, None, None, loc=loc, bc=bc, ext=1, coords=coords)
    set_both(g, o)
    c = 1e-5
    e = g + c * lap(g)
    for _ in range(4):
        g.assign(e)
        expect_tma()
        oracle.assign(o, e.signature(), [o] * 4, [c])
    a, b = gpu_storage(g, o)
    assert_same(a, b, exact, what=f"tma stretched loc={loc}")


@pytest.mark.parametrize("mode,exact", MODES)
def test_two_fields_not_aliased(engine, oracle, mode, exact):
    """u = v + c*lap(w): three distinct fields, four TMA slots (NS > 1), destination not among the operands (no twin buffer)"""
    host.set_mode(mode)
    dims = [131, 23, 19]
    bc = {(d, s): (capi.BC_NEUM, 0.0) for d in range(3) for s in range(2)}
    u, ou = make_pair(dims, [0] * 3, [1] * 3, bc=bc, ext=1, name="u")
    v, ov = make_pair_on(u.mesh, ou.mesh, bc=bc, ext=1, name="v")
    w, ow = make_pair_on(u.mesh, ou.mesh, bc=bc, ext=1, name="w")
    rng = np.random.default_rng(7)
    shape = u.localRange.shape(3)
    for gf, of in ((u, ou), (v, ov), (w, ow)):
        set_both(gf, of, arr=rng.standard_normal(shape))
    c = 3e-5
    e = v + c * lap(w)
    u.assign(e)
    expect_tma()
    oracle.assign(ou, e.signature(), [ov, ow, ow, ow], [c])
    a, b = gpu_storage(u, ou)
    assert_same(a, b, exact, what="tma two fields")


@pytest.mark.parametrize("mode,exact", MODES)
@pytest.mark.parametrize("op", [capi.OP_ADD, capi.OP_MINUS, capi.OP_MUL, capi.OP_DIV])
def test_compound_ops(engine, oracle, mode, exact, op):
    """u op= lap(v) (HASOP variant: the old destination value is read back through the `old` pointer)"""
    host.set_mode(mode)
    dims = [99, 21, 18]
    bc = {(d, s): (capi.BC_NEUM, 0.0) for d in range(3) for s in range(2)}
    u, ou = make_pair(dims, [0] * 3, [2, 1, 1], bc=bc, ext=1, name="u")
    v, ov = make_pair_on(u.mesh, ou.mesh, bc=bc, ext=1, name="v")
    rng = np.random.default_rng(11)
    shape = u.localRange.shape(3)
    set_both(u, ou, arr=rng.standard_normal(shape) + 3.0)
    # v = |x|^2 / 2 + small noise: lap(v) = 3 + O(0.1), bounded away from zero, so that u /= lap(v) is well conditioned
    # (dividing by a Laplacian of white noise amplifies the FAST-mode rounding differences past any fixed tolerance)
    lr = v.localRange
    ax = [np.linspace(0., hi, n)[lr.start[d]:lr.end[d]] for d, (n, hi) in enumerate(zip(dims, [2, 1, 1]))]
    smooth = 0.5 * (ax[0][:, None, None] ** 2 + ax[1][None, :, None] ** 2 + ax[2][None, None, :] ** 2)
    set_both(v, ov, arr=np.asfortranarray(smooth + 1e-7 * rng.standard_normal(shape)))
    e = lap(v)
    u.assign(e, op)
    expect_tma()
    oracle.assign(ou, e.signature(), [ov] * 3, [], op=op)
    a, b = gpu_storage(u, ou)
    assert_same(a, b, exact, what=f"tma compound op {op}")


@pytest.mark.parametrize("mode,exact", MODES)
def test_periodic_pad2(engine, oracle, mode, exact):
    """fully periodic cell-centred box, ext 2 / padding 2 (config C5's field set-up, TGMPI.cpp:35): ghost copies on every axis"""
    host.set_mode(mode)
    dims = [129, 25, 21]
    bc = {(d, s): (capi.BC_PERIODIC, 0.0) for d in range(3) for s in range(2)}
    g, o = make_pair(dims, [0] * 3, [2 * np.pi] * 3, loc=[1, 1, 1], bc=bc, ext=2, padding=2)
    set_both(g, o)
    c = 1e-4
    e = g + c * lap(g)
    for _ in range(4):
        g.assign(e)
        expect_tma()
        oracle.assign(o, e.signature(), [o] * 4, [c])
    a, b = gpu_storage(g, o)
    assert_same(a, b, exact, what="tma periodic pad 2")


@pytest.mark.parametrize("mode,exact", MODES)
def test_staggered_gradient_z(engine, oracle, mode, exact):
    """w = w - dt*dz(p) (LidDriven3D.cpp:86): two slots with different tap sets and locations (w: z-face, p: cell centre)"""
    host.set_mode(mode)
    dims = [97, 20, 33]
    per = {(d, s): (capi.BC_PERIODIC, 0.0) for d in range(3) for s in range(2)}
    w, ow = make_pair(dims, [0] * 3, [1] * 3, loc=[1, 1, 0], bc=per, ext=1, name="w")
    p, op_ = make_pair_on(w.mesh, ow.mesh, loc=[1, 1, 1], bc=per, ext=1, name="p")
    rng = np.random.default_rng(5)
    set_both(w, ow, arr=rng.standard_normal(w.localRange.shape(3)))
    set_both(p, op_, arr=rng.standard_normal(p.localRange.shape(3)))
    dt = 1e-3
    e = w - dt * dz(D1, p)
    w.assign(e)
    expect_tma()
    oracle.assign(ow, e.signature(), [ow, op_], [dt])
    a, b = gpu_storage(w, ow)
    assert_same(a, b, exact, what="tma staggered dz")


@pytest.mark.parametrize("mode", [capi.MODE_EXACT, capi.MODE_FAST])
def test_tma_vs_window_ab(engine, mode):
    """the same assignment through the TMA skeleton and through the register-window skeleton (option "tma" = 0): bit-equal in both
    arithmetic modes (both evaluate the identical functor; only the staging differs)"""
    host.set_mode(mode)
    l = capi.lib()
    n = 161
    res = {}
    try:
        for tma in (1, 0):
            capi.check(l.opf_set_option(b"tma", tma))
            g, o = make_pair([n, 45, 37], [0] * 3, [1] * 3, bc=dirc(3))
            set_both(g, o)
            e = g + 1e-5 * lap(g)
            for _ in range(3):
                g.assign(e)
            assert kernel_name() == ("opf::tma_kernel" if tma else "opf::window_kernel")
            res[tma] = g.to_numpy()
    finally:
        capi.check(l.opf_set_option(b"tma", 1))
    assert np.array_equal(res[0], res[1])


@pytest.mark.parametrize("n", [513])
def test_c2_full_size_against_the_reference(engine, n):
    """BASELINE config C2 at full size against the UNMODIFIED reference (oracle/_ref/bin/ref_explicit, built by oracle/build_ref.sh;
    the binary travels with the snapshot): 513^3 nodes, sin initial condition, 3 steps.  EXACT: every bit of all 513^3 values equal;
    FAST: <= 1e-12 relative L-inf."""
    if not O.ref_available("ref_explicit"):
        pytest.skip("oracle/_ref/bin/ref_explicit not built (needs /root/reference in the authoring container)")
    steps = 3
    with tempfile.TemporaryDirectory() as td:
        dump = os.path.join(td, "c2.opfd")
        O.run_ref("ref_explicit", "--case", "ftcs3d", "--n", n, "--steps", steps, "--init", "sin", "--threads", os.cpu_count() or 8, "--dump", dump)
        s, e, ref = O.read_opfd(dump)
        ref = np.array(ref, order="F")
    assert tuple(e[d] - s[d] for d in range(3)) == (n, n, n)
    mb = host.MeshBuilder(3).newMesh(n, n, n)
    for d in range(3):
        mb.setMeshOfDim(d, 0., 1.)
    b = host.ExprBuilder().setName("u").setMesh(mb.build())
    for d in range(3):
        b.setBC(d, 0, host.BCType.Dirc, 1.).setBC(d, 1, host.BCType.Dirc, 1.)
    u = b.build()
    # initBy(sin(pi x) sin(pi y) sin(pi z)) over the assignable range (CartesianField.hpp:283-294), evaluated with libm's sin like the
    # reference's std::sin; the product is formed in the reference's order (sx*sy)*sz
    a = u.assignableRange
    x, _, _ = u.mesh.axis(0)
    _, ext = u.mesh.ranges()
    sins = []
    for d in range(3):
        xd = u.mesh.axis(d)[0]
        sins.append(np.array([math.sin(math.pi * xd[i - ext.start[d]]) for i in range(a.start[d], a.end[d])]))
    init = np.asfortranarray((sins[0][:, None, None] * sins[1][None, :, None]) * sins[2][None, None, :])
    c = 0.1 / (n - 1) ** 2 * 1.0
    expr = u + c * lap(u)
    for mode, exact in MODES:
        host.set_mode(mode)
        u.assign(0.0)  # fields start zeroed in the reference (PlainTensor value-initialises)
        u.from_numpy(init, a)
        for _ in range(steps):
            u.assign(expr)
            expect_tma()
        got = u.to_numpy(capi.Range.make(list(s), list(e)))
        if exact:
            nbad = int(np.count_nonzero(got != ref))
            assert nbad == 0, f"EXACT mode: {nbad} of {ref.size} values differ from the reference, max abs diff {np.abs(got - ref).max()}"
        else:
            err = np.abs(got - ref).max() / np.abs(ref).max()
            assert err <= 1e-12, f"FAST mode: relative L-inf error {err} vs the reference at 513^3"
        del got
