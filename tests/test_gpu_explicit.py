"""GPU parity tests proper: the CUDA path through the C ABI vs the oracle on identical seeded inputs.

EXACT mode must be bit-identical to the oracle (which is itself pinned bit-for-bit to the reference, see
test_oracle_pinned.py); FAST mode must agree within 1e-12 relative L-inf (BASELINE.json north_star tolerance)."""
import numpy as np
import pytest

from opflow_b200 import capi, host
from opflow_b200.host import (D1FirstOrderBiasedDownwind, D1FirstOrderBiasedUpwind, D1FirstOrderCentered,
                              D1WENO53Downwind, D1WENO53Upwind, D2SecondOrderCentered, d1IntpCenterToCorner,
                              d1IntpCornerToCenter, d2x, d2y, d2z, dx, dy)
from helpers import assert_same, dirc, gpu_storage, make_pair, set_both

pytestmark = pytest.mark.gpu

MODES = [(capi.MODE_EXACT, True), (capi.MODE_FAST, False)]


def stretched(n, a=0.0, b=1.0):
    s = np.linspace(0, 1, n)
    return a + (b - a) * (s + 0.15 * np.sin(2 * np.pi * s) / (2 * np.pi))


@pytest.mark.parametrize("mode,exact", MODES)
@pytest.mark.parametrize("n,steps", [(65, 100), (130, 7)])
def test_ftcs2d(engine, oracle, mode, exact, n, steps):
    """examples/FTCS2D/FTCS-OMP.cpp:26 -- BASELINE config C1 at oracle-friendly size"""
    host.set_mode(mode)
    g, o = make_pair([n, n], [0, 0], [1, 1], bc=dirc(2))
    g.assign(0.0), o.fill(0.0)
    c = 0.1 / (n - 1) ** 2 * 1.0
    e = g + c * (d2x(D2SecondOrderCentered, g) + d2y(D2SecondOrderCentered, g))
    sig = e.signature()
    for _ in range(steps):
        g.assign(e)
        oracle.assign(o, sig, [o, o, o], [c])
    a, b = gpu_storage(g, o)
    assert_same(a, b, exact, what=f"ftcs2d n={n}")


@pytest.mark.parametrize("mode,exact", MODES)
def test_ftcs3d(engine, oracle, mode, exact):
    """BASELINE config C2 (7-point FTCS) at 33^3 with a non-trivial initial field"""
    host.set_mode(mode)
    n = 33
    g, o = make_pair([n] * 3, [0] * 3, [1] * 3, bc=dirc(3))
    set_both(g, o)
    c = 0.1 / (n - 1) ** 2
    D = D2SecondOrderCentered
    e = g + c * (d2x(D, g) + d2y(D, g) + d2z(D, g))
    for _ in range(10):
        g.assign(e)
        oracle.assign(o, e.signature(), [o] * 4, [c])
    a, b = gpu_storage(g, o)
    assert_same(a, b, exact, what="ftcs3d")


@pytest.mark.parametrize("mode,exact", MODES)
def test_ftcs3d_ragged_nonuniform(engine, oracle, mode, exact):
    """ragged extents (no tile multiple) + stretched mesh + Center location + Neumann/Dirichlet mix + ext 1"""
    host.set_mode(mode)
    dims = [37, 21, 13]
    coords = [stretched(n) for n in dims]
    bc = {(0, 0): (capi.BC_DIRC, 1.0), (0, 1): (capi.BC_NEUM, 0.5), (1, 0): (capi.BC_NEUM, -0.25), (1, 1): (capi.BC_DIRC, 2.0),
          (2, 0): (capi.BC_SYMM, 0.0), (2, 1): (capi.BC_ASYMM, 0.0)}
    g, o = make_pair(dims, None, None, loc=[1, 1, 1], bc=bc, ext=1, coords=coords)
    set_both(g, o)
    c = 1e-5
    D = D2SecondOrderCentered
    e = g + c * (d2x(D, g) + d2y(D, g) + d2z(D, g))
    for _ in range(5):
        g.assign(e)
        oracle.assign(o, e.signature(), [o] * 4, [c])
    a, b = gpu_storage(g, o)
    assert_same(a, b, exact, what="ftcs3d ragged")


@pytest.mark.parametrize("mode,exact", MODES)
@pytest.mark.parametrize("kind", ["weno_down", "weno_up", "upwind1"])
def test_conv1d(engine, oracle, mode, exact, kind):
    """examples/CONV1D/CONV1D.cpp:18-31 -- BASELINE config C3 (WENO5) at 2^12+1 nodes, top-hat initial condition"""
    host.set_mode(mode)
    n = 4097
    bc = {(0, 0): (capi.BC_DIRC, 0.0), (0, 1): (capi.BC_DIRC, 0.0)}
    g, o = make_pair([n], [0], [1], bc=bc, ext=3)
    tophat = lambda x: 1.0 if 0.2 <= x[0] <= 0.4 else 0.0
    g.initBy(tophat), o.init_by(tophat)
    c = 0.5 / (n - 1) * 1.0
    e = {"weno_down": g - c * dx(D1WENO53Downwind, g), "weno_up": g + c * dx(D1WENO53Upwind, g),
         "upwind1": g - c * dx(D1FirstOrderBiasedDownwind, g)}[kind]
    for _ in range(40):
        g.assign(e)
        oracle.assign(o, e.signature(), [o, o], [c])
    a, b = gpu_storage(g, o)
    assert_same(a, b, exact, tol=1e-11 if kind.startswith("weno") else 1e-12, what=kind)


@pytest.mark.parametrize("mode,exact", MODES)
def test_single_operators(engine, oracle, mode, exact):
    """every stencil Op::eval on both locations, on a stretched 2-D mesh"""
    host.set_mode(mode)
    dims = [45, 29]
    coords = [stretched(n, 0, 2) for n in dims]
    for loc in ([0, 0], [1, 1], [0, 1], [1, 0]):
        bc = {(d, s): (capi.BC_NEUM, 0.0) for d in range(2) for s in range(2)}
        g, o = make_pair(dims, None, None, loc=loc, bc=bc, ext=3, coords=coords, name="src")
        set_both(g, o)
        ops = []
        for ax in (0, 1):
            ops += [d2x(D2SecondOrderCentered, g) if ax == 0 else d2y(D2SecondOrderCentered, g),
                    host.d1(D1FirstOrderCentered, ax, g), host.d1(D1FirstOrderBiasedDownwind, ax, g),
                    host.d1(D1FirstOrderBiasedUpwind, ax, g), host.d1(D1WENO53Downwind, ax, g), host.d1(D1WENO53Upwind, ax, g)]
            ops.append(d1IntpCenterToCorner(ax, g) if loc[ax] == 1 else d1IntpCornerToCenter(ax, g))
        for e in ops:
            sig = e.signature()
            r, eloc = host.prepared(e, capi.R_ACCESSIBLE)
            ro, oloc = oracle.prepare(sig, [o], 2)
            assert r.tup(2) == ro.tup(2) and eloc[:2] == oloc[:2], sig
            # destination with the expression's loc, no BC, big enough ext to hold everything
            gd, od = make_pair(dims, None, None, loc=eloc[:2], ext=0, coords=coords, name="dst")
            gd.assign(e)
            oracle.assign(od, sig, [o], [])
            # compare where the expression is defined
            lo = [max(r.start[d], gd.localRange.start[d]) for d in range(2)]
            hi = [min(r.end[d], gd.localRange.end[d]) for d in range(2)]
            rr = capi.Range.make(lo, hi)
            a, b = gd.to_numpy(rr), od.view((lo, hi))
            assert_same(a, b, exact, tol=1e-11 if "Weno" in sig else 1e-12, what=f"{sig} loc={loc}")


@pytest.mark.parametrize("mode,exact", MODES)
def test_liddriven_explicit_updates(engine, oracle, mode, exact):
    """examples/LidDriven/LidDriven2D.cpp:83-90: staggered explicit updates (conv_xy, dx/dy of dp, Poisson rhs)"""
    host.set_mode(mode)
    n = 33
    dims = [n, n]
    D0 = {(d, s): (capi.BC_DIRC, 0.0) for d in range(2) for s in range(2)}
    lid = dict(D0)
    lid[(1, 1)] = (capi.BC_DIRC, 1.0)
    NE = {(d, s): (capi.BC_NEUM, 0.0) for d in range(2) for s in range(2)}
    u, uo = make_pair(dims, [0, 0], [1, 1], loc=[0, 1], bc=lid, ext=1, name="u")
    du, duo = make_pair(dims, [0, 0], [1, 1], loc=[0, 1], bc=D0, ext=1, name="du")
    v, vo = make_pair(dims, [0, 0], [1, 1], loc=[1, 0], bc=D0, ext=1, name="v")
    dv, dvo = make_pair(dims, [0, 0], [1, 1], loc=[1, 0], bc=D0, ext=1, name="dv")
    p, po = make_pair(dims, [0, 0], [1, 1], loc=[1, 1], bc=NE, ext=1, name="p")
    dp, dpo = make_pair(dims, [0, 0], [1, 1], loc=[1, 1], bc=NE, ext=1, name="dp")
    for gf, of in ((u, uo), (du, duo), (v, vo), (dv, dvo), (p, po), (dp, dpo)):
        set_both(gf, of)
    dt = 0.5e-2
    C = D1FirstOrderCentered
    conv_xy = lambda a, b: dy(C, d1IntpCenterToCorner(1, a) * d1IntpCenterToCorner(0, b))
    steps = [
        (du, duo, du - (0.5 * dt) * conv_xy(u, dv), [duo, uo, dvo], [0.5 * dt]),
        (u, uo, u + du, [uo, duo], []),
        (v, vo, v + dv, [vo, dvo], []),
        (p, po, (dx(C, du) + dy(C, dv)) / dt, [duo, dvo], [dt]),
        (u, uo, u - dt * dx(C, dp), [uo, dpo], [dt]),
        (v, vo, v - dt * dy(C, dp), [vo, dpo], [dt]),
        (p, po, p + dp, [po, dpo], []),
    ]
    for gd, od, e, ofs, sc in steps:
        gd.assign(e)
        oracle.assign(od, e.signature(), ofs, sc)
        a, b = gpu_storage(gd, od)
        assert_same(a, b, exact, what=e.signature())


def test_compound_and_scalar_assign(engine, oracle):
    """Expr::operator+=,-=,*=,/= (Expr.hpp:59-117) and assignImpl_final(const D&) (CartesianField.hpp:237-280)"""
    host.set_mode(capi.MODE_EXACT)
    g, o = make_pair([40, 23], [0, 0], [1, 2], bc=dirc(2, 0.5), ext=1)
    h, ho = make_pair([40, 23], [0, 0], [1, 2], bc=dirc(2, 0.25), ext=1, name="h")
    set_both(g, o), set_both(h, ho, arr=np.random.default_rng(7).uniform(0.5, 2.0, [40, 23]))
    for op, c in ((capi.OP_ADD, 0.75), (capi.OP_MINUS, 0.125), (capi.OP_MUL, 1.5), (capi.OP_DIV, 3.0), (capi.OP_EQ, 2.0)):
        g.assign(c, op)
        oracle.assign(o, "S<0>", [], [c], op)
        a, b = gpu_storage(g, o)
        assert_same(a, b, True, what=f"scalar op {op}")
    set_both(g, o)
    for op in (capi.OP_ADD, capi.OP_MINUS, capi.OP_MUL, capi.OP_DIV):
        g.assign(h, op)
        oracle.assign(o, "F<0>", [ho], [], op)
        a, b = gpu_storage(g, o)
        assert_same(a, b, True, what=f"field op {op}")
    # aliased compound with a stencil: u -= c*d2x(u) spelled through a builtin (u + c*d2x(u))
    e = g + 0.001 * d2x(D2SecondOrderCentered, g)
    g.assign(e)
    oracle.assign(o, e.signature(), [o, o], [0.001])
    a, b = gpu_storage(g, o)
    assert_same(a, b, True, what="aliased 1-D stencil in a 2-D field")


def test_unregistered_expression_fails_loudly(engine):
    g, _ = make_pair([16, 16], [0, 0], [1, 1], bc=dirc(2))
    e = host.sqrt(host.abs_(g)) * 3.0 + d2y(D2SecondOrderCentered, g)
    with pytest.raises(capi.EngineError, match="no compiled device kernel"):
        g.assign(e)


def test_out_of_storage_read_is_refused(engine):
    """WENO on a field without ext would read outside the storage: OPF_ERR_RANGE instead of UB"""
    bc = {(0, 0): (capi.BC_DIRC, 0.0), (0, 1): (capi.BC_DIRC, 0.0)}
    g, _ = make_pair([64], [0], [1], bc=bc, ext=0)
    with pytest.raises(capi.EngineError, match="outside its storage"):
        g.assign(g - 0.1 * dx(D1WENO53Downwind, g))


@pytest.mark.parametrize("dims", [[1000], [123, 77], [31, 17, 23]])
def test_reductions(engine, oracle, dims):
    """rangeReduce (RangeFor.hpp:87-121): sum / max / min / abs-max / sum of squares vs numpy on the same values"""
    dim = len(dims)
    g, o = make_pair(dims, [0] * dim, [1] * dim, bc=dirc(dim, 0.0))
    set_both(g, o)
    ref = o.local()
    r = g.localRange
    got = {k: host.rangeReduce(g, k, r) for k in range(5)}
    assert abs(got[capi.RED_SUM] - ref.sum()) <= 1e-12 * np.abs(ref).sum()
    assert got[capi.RED_MAX] == ref.max() and got[capi.RED_MIN] == ref.min()
    assert got[capi.RED_ABSMAX] == np.abs(ref).max()
    assert abs(got[capi.RED_SUMSQ] - (ref ** 2).sum()) <= 1e-12 * (ref ** 2).sum()
    # determinism: same launch twice -> bitwise same
    assert host.rangeReduce(g, capi.RED_SUM, r) == got[capi.RED_SUM]
    # dot product through an expression
    d = host.rangeReduce(g * g, capi.RED_SUM, r)
    assert abs(d - (ref * ref).sum()) <= 1e-12 * (ref * ref).sum()


@pytest.mark.parametrize("mode", [capi.MODE_EXACT, capi.MODE_FAST])
@pytest.mark.parametrize("dims,count", [((129, 97), 100), ((65, 33, 17), 71), ((257,), 40)])
def test_assign_repeat_equals_plain_assignments(engine, mode, dims, count):
    """opf_assign_repeat (the time loop replayed from CUDA graphs of 32 steps) == the same number of plain assignments, bit for bit:
    ping-pong buffers, BC fills and the remainder steps included; called twice so that the second call replays a cached graph"""
    import ctypes as C
    host.set_mode(mode)
    dim = len(dims)
    res = []
    for use_repeat in (False, True):
        g, _ = make_pair(list(dims), [0] * dim, [1] * dim, bc=dirc(dim))
        rng = np.random.default_rng(9)
        g.from_numpy(np.asfortranarray(rng.standard_normal(g.localRange.shape(dim))))
        c = 0.05 / (max(dims) - 1) ** 2
        lap = d2x(D2SecondOrderCentered, g)
        if dim >= 2:
            lap = lap + d2y(D2SecondOrderCentered, g)
        if dim == 3:
            lap = lap + d2z(D2SecondOrderCentered, g)
        e = g + c * lap
        sig, fields, scalars = e.flatten()
        F = (C.c_void_p * len(fields))(*[f.h for f in fields])
        S = (C.c_double * len(scalars))(*scalars)
        for _ in range(2):
            if use_repeat:
                capi.check(engine.opf_assign_repeat(g.h, capi.OP_EQ, sig.encode(), F, len(fields), S, len(scalars), count))
            else:
                for _ in range(count):
                    g.assign(e)
        res.append(g.to_numpy(g.getLocalReadableRange()))
    assert np.array_equal(res[0], res[1])
