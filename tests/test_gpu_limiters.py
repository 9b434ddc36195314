"""Flux-limiter interpolators and Convolution on the device against (a) fixtures computed by the UNMODIFIED reference
(tests/golden/ref_ops.json) and (b) the oracle on a larger random case.  Reference: D1FluxLimiter.hpp:41-203,
FluxLimiterKernels.hpp:32-80, Convolution.hpp:45-82.  EXACT mode bit-identical, FAST within 1e-12 relative."""
import numpy as np
import pytest

import ops_golden as G
from opflow_b200 import capi, host
from helpers import assert_same, make_pair, make_pair_on, set_both

pytestmark = pytest.mark.gpu
MODES = [(capi.MODE_EXACT, True), (capi.MODE_FAST, False)]


def build_gpu(mesh, loc, fn):
    b = host.ExprBuilder().setMesh(mesh).setLoc(loc).setExt(2)
    for d in range(2):
        b.setBC(d, 0, host.BCType.Neum, 0.).setBC(d, 1, host.BCType.Neum, 0.)
    f = b.build()
    f.initBy(fn)
    return f


def expression(case, fields):
    if case["node"].startswith("Conv"):
        return host.conv(fields[0], G.kernel_of(case["node"]))
    return host.d1IntpFl(case["node"], case["axis"], case["dir"], fields[0], fields[1])


@pytest.mark.parametrize("mode,exact", MODES)
def test_reference_fixtures(engine, mode, exact):
    host.set_mode(mode)
    d = G.load()
    cx, cy = G.coords(d["nx"], d["ny"])
    mesh = host.MeshBuilder(2).newMesh(d["nx"], d["ny"]).setMeshOfDim(0, cx).setMeshOfDim(1, cy).build()
    for c in d["cases"]:
        sig, leaves, scal, _ = G.describe(c)
        fields = [build_gpu(mesh, loc, fn) for loc, fn in leaves]
        e = expression(c, fields)
        assert e.signature() == sig
        r, loc = host.prepared(e, capi.R_ACCESSIBLE)
        assert [list(x) for x in r.tup(2)] == c["acc"] and loc[:2] == c["loc"], sig
        assert [list(x) for x in host.prepared(e, capi.R_LOCAL)[0].tup(2)] == c["local"], sig
        assert [list(x) for x in host.prepared(e, capi.R_LOGICAL)[0].tup(2)] == c["logical"], sig
        # destination: a field of the result's location with room for the whole accessible range
        dst = host.ExprBuilder().setMesh(mesh).setLoc(c["loc"]).setExt(2).build()
        dst.assign(0.0)
        dst.assign(e)
        w = dst.assignableRange
        lo = [max(c["acc"][0][k], w.start[k]) for k in range(2)]
        hi = [min(c["acc"][1][k], w.end[k]) for k in range(2)]
        got = dst.to_numpy(capi.Range.make(lo, hi))
        ref = G.reference_values(c)[lo[0] - c["acc"][0][0]:hi[0] - c["acc"][0][0], lo[1] - c["acc"][0][1]:hi[1] - c["acc"][0][1]]
        assert got.size > 0
        assert_same(got, ref, exact, what=sig)


@pytest.mark.parametrize("mode,exact", MODES)
@pytest.mark.parametrize("scheme", host.FLUX_LIMITERS)
def test_against_oracle_random(engine, oracle, mode, exact, scheme):
    """larger stretched mesh, random data (sign changes of the advecting field everywhere), both directions"""
    host.set_mode(mode)
    dims = [75, 41]
    s = [np.linspace(0, 1, n) for n in dims]
    coords = [1.5 * (s[0] + 0.1 * np.sin(2 * np.pi * s[0]) / (2 * np.pi)), s[1] ** 1.3]
    bc = {(d, k): (capi.BC_NEUM, 0.0) for d in range(2) for k in range(2)}
    rng = np.random.default_rng(17)
    for direction in ("C2N", "N2C"):
        le, lu = ([1, 1], [0, 1]) if direction == "C2N" else ([0, 1], [1, 1])
        e, oe = make_pair(dims, None, None, loc=le, bc=bc, ext=2, coords=coords, name="e")
        u, ou = make_pair_on(e.mesh, oe.mesh, loc=lu, bc=bc, ext=2, name="u")
        dst, od = make_pair_on(e.mesh, oe.mesh, loc=lu, ext=0, name="dst")
        set_both(e, oe, arr=rng.standard_normal(e.localRange.shape(2)))
        set_both(u, ou, arr=rng.standard_normal(u.localRange.shape(2)))
        ex = host.d1IntpFl(scheme, 0, direction, u, e)
        dst.assign(0.0), od.fill(0.0)
        dst.assign(ex)
        oracle.assign(od, ex.signature(), [ou, oe], [])
        r = dst.assignableRange
        # Harmonic / Albada divide by r + 1 resp. r^2 + 1 with r = slope ratio of random data: wide dynamic range -> compare per
        # cell against the cell's own scale in FAST mode
        a, b = dst.to_numpy(r), od.view(r.tup(2))
        # Harmonic is 0 / 0 where the two slopes cancel exactly (r = -1; it happens in mirrored Neumann ghost rows): the reference
        # produces NaN there and so must the device, at the same cells
        assert np.array_equal(np.isnan(a), np.isnan(b)), f"{scheme} {direction}: NaN pattern differs"
        ok = ~np.isnan(b)
        if exact:
            assert np.array_equal(a[ok], b[ok]), f"{scheme} {direction}"
        else:
            assert (np.abs(a[ok] - b[ok]) <= 1e-12 * np.maximum(np.abs(b[ok]), 1.0)).all(), f"{scheme} {direction}: {np.abs(a[ok] - b[ok]).max()}"


@pytest.mark.parametrize("mode,exact", MODES)
def test_convolution_3d_and_composed(engine, oracle, mode, exact):
    """3 x 3 x 3 kernel (27 scalar slots) on a 3-D field, and h*h*conv(a*b, k) -- the integral operator of UniLS.cpp:113"""
    host.set_mode(mode)
    rng = np.random.default_rng(3)
    bc = {(d, k): (capi.BC_NEUM, 0.0) for d in range(3) for k in range(2)}
    u, ou = make_pair([23, 17, 12], [0] * 3, [1] * 3, loc=[1, 1, 1], bc=bc, ext=1, name="u")
    d3, od3 = make_pair_on(u.mesh, ou.mesh, loc=[1, 1, 1], ext=0, name="d")
    set_both(u, ou, arr=rng.standard_normal(u.localRange.shape(3)))
    k3 = rng.standard_normal((3, 3, 3))
    e = host.conv(u, k3)
    assert e.signature() == "Conv<3,3,3,0,F<0>>"
    d3.assign(e)
    _, _, scal = e.flatten()
    oracle.assign(od3, e.signature(), [ou], scal)
    r = d3.assignableRange
    assert_same(d3.to_numpy(r), od3.view(r.tup(3)), exact, what="conv 3x3x3")
    bc2 = {(d, k): (capi.BC_NEUM, 0.0) for d in range(2) for k in range(2)}
    a, oa = make_pair([40, 31], [0, 0], [1, 1], loc=[1, 1], bc=bc2, ext=1, name="a")
    b, ob = make_pair_on(a.mesh, oa.mesh, loc=[1, 1], bc=bc2, ext=1, name="b")
    d2, od2 = make_pair_on(a.mesh, oa.mesh, loc=[1, 1], ext=0, name="d")
    set_both(a, oa, arr=rng.standard_normal(a.localRange.shape(2)))
    set_both(b, ob, arr=rng.standard_normal(b.localRange.shape(2)))
    h = 1. / 39
    e2 = (h * h) * host.conv(a * b, G.kernel_of("Conv33"))
    assert e2.signature() == "Mul<S<0>,Conv<3,3,1,1,Mul<F<0>,F<1>>>>"
    d2.assign(e2)
    _, _, scal = e2.flatten()
    oracle.assign(od2, e2.signature(), [oa, ob], scal)
    r = d2.assignableRange
    assert_same(d2.to_numpy(r), od2.view(r.tup(2)), exact, what="h*h*conv(a*b)")
