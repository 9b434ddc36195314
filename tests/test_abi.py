"""CPU-side checks of the product library: it loads, exports every symbol include/opflow_b200.h declares, its host logic
(mesh arrays, split strategies) is bit-identical to the oracle/reference, and it refuses to compute without a GPU."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from opflow_b200 import capi, host
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_fields.json")))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "opflow_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(opf_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"opf_expr_launcher"}
    assert len(declared) >= 45
    l = capi.lib()
    for name in sorted(declared):
        assert hasattr(l, name), f"libopflow_b200.so does not export {name}"
    assert declared <= set(capi.SIGNATURES), declared - set(capi.SIGNATURES)


def test_builtin_expressions_registered():
    l = capi.lib()
    names = {l.opf_expr_builtin_name(i).decode() for i in range(l.opf_expr_builtin_count())}
    for sig in ("Add<F<0>,Mul<S<0>,Add<D2C<0,F<1>>,D2C<1,F<2>>>>>", "Add<F<0>,Mul<S<0>,Add<Add<D2C<0,F<1>>,D2C<1,F<2>>>,D2C<2,F<3>>>>>",
                "Sub<F<0>,Mul<S<0>,WenoDn<0,F<1>>>>", "S<0>", "F<0>"):
        assert sig in names
        assert l.opf_expr_is_registered(sig.encode())


@pytest.mark.skipif(capi.lib().opf_device_count() > 0, reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    m = host.MeshBuilder(2).newMesh(9, 9).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).build()
    with pytest.raises(capi.EngineError, match="no CUDA device"):
        host.ExprBuilder().setMesh(m).build()


def stretched(n, scale=1.0):
    s = np.arange(n).astype(np.float64) / (n - 1)
    return scale * (s + 0.15 * np.sin(2 * np.pi * s) / (2 * np.pi))


@pytest.mark.parametrize("mode", [capi.MESHEXT_UNDEFINED, capi.MESHEXT_SYMM, capi.MESHEXT_PERIODIC, capi.MESHEXT_UNIFORM])
def test_mesh_arrays_bit_identical_to_oracle(mode):
    """MeshBuilder::set1DMesh / setExtMesh (CartesianMesh.hpp:208-303): x, dx, idx over the 5-cell extended range"""
    for n, lo, hi, coords in ((11, 0.0, 2.0, None), (1025, 0.0, 1.0, None), (17, -1.0, 3.0, None), (33, None, None, stretched(33, 2.0))):
        mb = host.MeshBuilder(1).newMesh(n).setExtMode(mode)
        mb.setMeshOfDim(0, coords) if coords is not None else mb.setMeshOfDim(0, lo, hi)
        g = mb.build()
        x, dx, idx = g.axis(0)
        om = O.Mesh([n], [lo], [hi], coords=[coords] if coords is not None else None, ext_mode=[mode])
        assert np.array_equal(x, om.x[0]) and np.array_equal(dx, om.dx[0])
        if mode != capi.MESHEXT_UNIFORM:  # reference quirk: Uniform mode computes 1/idx of an unset entry -> inf/NaN
            assert np.array_equal(idx, om.idx[0])


def test_split_even_bit_identical_to_reference():
    l = capi.lib()
    for c in REF["split"]:
        dim, p = c["dim"], c["ranks"]
        r = capi.Range.make([0] * dim, c["mesh"])
        out = (capi.Range * p)()
        capi.check(l.opf_split_even(dim, C.byref(r), p, out))
        got = [[list(o.tup(dim)[0]), list(o.tup(dim)[1])] for o in out]
        assert got == c["map"], c


def test_split_slab():
    l = capi.lib()
    r = capi.Range.make([0, 0, 0], [1025, 1025, 1025])
    out = (capi.Range * 8)()
    capi.check(l.opf_split_slab(3, C.byref(r), 8, out))
    assert [o.tup() for o in out] == [((0, 0, 128 * k), (1024, 1024, 128 * (k + 1))) for k in range(8)]
    out3 = (capi.Range * 3)()
    capi.check(l.opf_split_slab(3, C.byref(capi.Range.make([0, 0, 0], [9, 9, 12])), 3, out3))
    assert [o.tup()[0][2] for o in out3] == [0, 3, 6] and out3[2].end[2] == 11


def test_signature_generation_matches_grammar():
    class Dummy(host.Field):
        def __init__(self):
            host.Expr.__init__(self, "F")
            self.h = None

        def __del__(self):
            pass
    u = Dummy()
    D2 = host.D2SecondOrderCentered
    e = u + 0.1 * (host.d2x(D2, u) + host.d2y(D2, u))
    sig, fields, scalars = e.flatten()
    assert sig == "Add<F<0>,Mul<S<0>,Add<D2C<0,F<1>>,D2C<1,F<2>>>>>" and len(fields) == 3 and scalars == [0.1]
    e = u - 0.5 * host.dx(host.D1WENO53Downwind, u)
    assert e.signature() == "Sub<F<0>,Mul<S<0>,WenoDn<0,F<1>>>>"


# ---- host logic through the real C ABI, on the CPU: opf_field_plan needs no device --------------------------------------------
def _plan(dims, lo, hi, coords, loc, bc, ext, split=None):
    dim = len(dims)
    mb = host.MeshBuilder(dim).newMesh(*dims)
    for d in range(dim):
        mb.setMeshOfDim(d, coords[d]) if coords is not None else mb.setMeshOfDim(d, lo[d], hi[d])
    b = host.ExprBuilder().setMesh(mb.build()).setLoc(loc).setExt(ext)
    for (d, s), (t, v) in bc.items():
        if t != capi.BC_UNDEFINED:
            b.setBC(d, s, t, v)
    if split:
        b.setPadding(max(1, ext)).setSplitStrategy(*split)
    return b.plan()


def _rng(r, dim):
    return [list(r.tup(dim)[0]), list(r.tup(dim)[1])]


def test_range_classification_table_bit_exact_on_cpu():
    """every row of the reference's range table (ExprBuilder::calculateRanges, CartesianField.hpp:950-1003) as dumped by the
    unmodified reference (tests/golden/ref_fields.json: loc x BC start x BC end x ext x uniform/stretched mesh)"""
    from test_oracle_pinned import BC, stretched
    n_checked = 0
    for c in REF["ranges1d"]:
        n = c["n"]
        bc = {(0, 0): (BC[c["bc"][0]], c["bcv"][0]), (0, 1): (BC[c["bc"][1]], c["bcv"][1])}
        u = _plan([n], [0.0], [2.0], [stretched(n)] if c["stretched"] else None, c["loc"], bc, c["ext"])
        r = c["ranges"]
        assert _rng(u.localRange, 1) == r["local"] and _rng(u.assignableRange, 1) == r["assignable"], c
        assert _rng(u.accessibleRange, 1) == r["accessible"] and _rng(u.logicalRange, 1) == r["logical"] and u.padding == r["padding"], c
        n_checked += 1
    assert n_checked == len(REF["ranges1d"]) and n_checked >= 100
    for c in REF["ghost2d"]:
        nx, ny = c["dims"]
        bc = {}
        for d in range(2):
            t0, v0, t1, v1 = c["bc"][d]
            bc[(d, 0)], bc[(d, 1)] = (BC[t0], float(v0)), (BC[t1], float(v1))
        u = _plan([nx, ny], [0.0, 0.0], [2.0, 1.0], None, c["loc"], bc, c["ext"])
        for key, got in (("local", u.localRange), ("assignable", u.assignableRange), ("accessible", u.accessibleRange), ("logical", u.logicalRange)):
            assert _rng(got, 2) == c["ranges"][key], (c["loc"], c["bc"], key)


def test_prepared_expression_ranges_bit_exact_on_cpu():
    """Expr::prepare() range algebra and result location of every stencil operator and of the FTCS / WENO expressions
    (opf_expr_prepare on plan-only fields) against the unmodified reference's dump"""
    l = capi.lib()
    bc = {(0, 0): (capi.BC_DIRC, 1.0), (0, 1): (capi.BC_NEUM, 0.0), (1, 0): (capi.BC_NEUM, 0.0), (1, 1): (capi.BC_DIRC, 0.0)}
    for c in REF["prepare2d"]:
        u = _plan([12, 10], [0.0, 0.0], [2.0, 1.0], None, c["loc"], bc, 3)
        nf = c["sig"].count("F<")
        F = (C.c_void_p * nf)(*[u.h] * nf)
        for which, key in ((capi.R_ACCESSIBLE, "acc"), (capi.R_LOCAL, "local"), (capi.R_LOGICAL, "logical")):
            r, loc = capi.Range(), (C.c_int * 3)()
            capi.check(l.opf_expr_prepare(c["sig"].encode(), F, nf, which, C.byref(r), loc))
            assert _rng(r, 2) == c[key] and list(loc)[:2] == c["eloc"], (c["sig"], key)


def test_expression_errors_are_reported_not_swallowed():
    """operands at different mesh locations abort in the reference (BinOpDefMacros.hpp.in:25-42): here OPF_ERR_LOC; malformed and
    unknown signatures are OPF_ERR_INVALID; a plan-only field refuses device work"""
    l = capi.lib()
    bc = {(d, s): (capi.BC_DIRC, 0.0) for d in range(2) for s in range(2)}
    a = _plan([9, 9], [0., 0.], [1., 1.], None, [0, 0], bc, 1)
    b = _plan([9, 9], [0., 0.], [1., 1.], None, [1, 0], bc, 1)
    F = (C.c_void_p * 2)(a.h, b.h)
    r = capi.Range()
    assert l.opf_expr_prepare(b"Add<F<0>,F<1>>", F, 2, capi.R_ACCESSIBLE, C.byref(r), None) == 6  # OPF_ERR_LOC
    assert b"loc" in l.opf_last_error()
    assert l.opf_expr_prepare(b"D1C<0,F<1>>", F, 2, capi.R_ACCESSIBLE, C.byref(r), None) == 0  # Center -> Corner on axis 0
    assert l.opf_expr_prepare(b"Add<F<0>,D1C<0,F<1>>>", F, 2, capi.R_ACCESSIBLE, C.byref(r), None) == 0  # now both Corner
    for bad in (b"Add<F<0>", b"Foo<F<0>>", b"D2C<7,F<0>>", b"Add<F<0>,F<1>>>"):
        assert l.opf_expr_prepare(bad, F, 2, capi.R_ACCESSIBLE, C.byref(r), None) == 2, bad  # OPF_ERR_INVALID
    rc = l.opf_field_update_padding(a.h)
    assert rc != 0  # no device here (OPF_ERR_NO_DEVICE) or, on a GPU box, "field is a plan"


def test_slab_and_even_split_neighbours_on_cpu():
    """updateNeighbors (CartesianField.hpp:298-347) through plan-only fields: a 4-rank slab split has 1 neighbour at the ends and 2
    inside; the reference's EvenSplitStrategy on 4 ranks in 2-D gives 2 x 2 blocks with 3 neighbours each (two edges, one corner);
    send boxes lie inside the local block, receive boxes outside it"""
    bc = {(d, s): (capi.BC_DIRC, 0.0) for d in range(3) for s in range(2)}
    mesh = host.MeshBuilder(3).newMesh(17, 17, 33).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).setMeshOfDim(2, 0., 2.).build()
    for rank, want in ((0, 1), (1, 2), (2, 2), (3, 1)):
        b = host.ExprBuilder().setMesh(mesh).setPadding(1)
        for (d, s), (t, v) in bc.items():
            b.setBC(d, s, t, v)
        u = b.setSplitStrategy(4, rank, host.split_slab(mesh, 4)).plan()
        nb = u.neighbors()
        assert len(nb) == want and all(abs(p - rank) == 1 for p, *_ in nb)
        lo, hi = u.localRange.tup(3)
        for p, send, recv, code in nb:
            assert all(lo[d] <= send[0][d] and send[1][d] <= hi[d] for d in range(3))
            assert recv[0][2] >= hi[2] or recv[1][2] <= lo[2]
    mesh2 = host.MeshBuilder(2).newMesh(33, 33).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).build()
    for rank in range(4):
        b = host.ExprBuilder().setMesh(mesh2).setLoc([1, 1]).setExt(1).setPadding(1)
        for d in range(2):
            for s in range(2):
                b.setBC(d, s, capi.BC_DIRC, 0.0)
        u = b.setSplitStrategy(4, rank, host.split_even(mesh2, 4)).plan()
        assert len(u.neighbors()) == 3
        sizes = sorted((s[1][0] - s[0][0]) * (s[1][1] - s[0][1]) for _, s, _, _ in u.neighbors())
        assert sizes == [1, 16, 16]  # one corner cell, two 16-cell edges


def test_flux_limiter_and_convolution_ranges_without_a_device():
    """opf_expr_prepare on plan-only fields (no GPU): accessible / local / logical ranges and result location of the flux-limiter
    interpolators and convolutions equal what the unmodified reference prepared (tests/golden/ref_ops.json; D1FluxLimiter.hpp:155-203,
    Convolution.hpp:66-82); a flux limiter applied to a field at the wrong location is refused like the reference's OP_ASSERT"""
    import ops_golden as G
    l = capi.lib()
    d = G.load()
    cx, cy = G.coords(d["nx"], d["ny"])
    mesh = host.MeshBuilder(2).newMesh(d["nx"], d["ny"]).setMeshOfDim(0, cx).setMeshOfDim(1, cy).build()

    def plan(loc):
        b = host.ExprBuilder().setMesh(mesh).setLoc(loc).setExt(2)
        for ax in range(2):
            b.setBC(ax, 0, host.BCType.Neum, 0.).setBC(ax, 1, host.BCType.Neum, 0.)
        return b.build(plan_only=True)

    for c in d["cases"]:
        sig, leaves, _, _ = G.describe(c)
        fields = [plan(loc) for loc, _ in leaves]
        F = (C.c_void_p * len(fields))(*[f.h for f in fields])
        for which, key in ((capi.R_ACCESSIBLE, "acc"), (capi.R_LOCAL, "local"), (capi.R_LOGICAL, "logical")):
            r, loc = capi.Range(), (C.c_int * 3)()
            capi.check(l.opf_expr_prepare(sig.encode(), F, len(fields), which, C.byref(r), loc))
            assert [list(x) for x in r.tup(2)] == c[key] and list(loc)[:2] == c["loc"], (sig, key)
    u, e = plan([0, 1]), plan([0, 1])  # e is Corner along x: Cen2Cor must refuse it
    F = (C.c_void_p * 2)(u.h, e.h)
    r, loc = capi.Range(), (C.c_int * 3)()
    assert l.opf_expr_prepare(b"FlQuickC2N<0,F<0>,F<1>>", F, 2, capi.R_ACCESSIBLE, C.byref(r), loc) != 0


def test_parity_leaf_of_the_red_black_smoother_without_a_device():
    """Par<c> (the colour mask of the red-black Gauss-Seidel half-sweeps, PFMG relaxType 2 / 3) is a scalar-like leaf: it is part of
    the grammar, its half-sweep kernels are compiled in for the Poisson operators, and it neither adds ranges nor field / scalar slots"""
    l = capi.lib()
    for sig in ("Mul<Par<0>,Mul<F<0>,F<1>>>", "Add<F<0>,Mul<Par<1>,Mul<F<1>,F<2>>>>",
                "Add<F<0>,Mul<Par<0>,Mul<F<1>,Sub<F<2>,Add<D2C<0,F<0>>,D2C<1,F<0>>>>>>>",
                "Add<F<0>,Mul<Par<1>,Mul<F<1>,Sub<F<2>,Add<Add<D2C<0,F<0>>,D2C<1,F<0>>>,D2C<2,F<0>>>>>>>"):
        assert l.opf_expr_is_registered(sig.encode()), sig
    mesh = host.MeshBuilder(2).newMesh(9, 7).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).build()
    b = host.ExprBuilder().setMesh(mesh).setLoc([1, 1]).setExt(1)
    for ax in range(2):
        b.setBC(ax, 0, host.BCType.Neum, 0.).setBC(ax, 1, host.BCType.Neum, 0.)
    x, dinv, rhs = (b.build(plan_only=True) for _ in range(3))
    F = (C.c_void_p * 3)(x.h, dinv.h, rhs.h)
    got = {}
    for sig in (b"Add<F<0>,Mul<Par<1>,Mul<F<1>,F<2>>>>", b"Add<F<0>,Mul<S<0>,Mul<F<1>,F<2>>>>"):
        r, loc = capi.Range(), (C.c_int * 3)()
        capi.check(l.opf_expr_prepare(sig, F, 3, capi.R_ACCESSIBLE, C.byref(r), loc))
        got[sig] = (r.tup(2), list(loc)[:2])
    assert got[b"Add<F<0>,Mul<Par<1>,Mul<F<1>,F<2>>>>"] == got[b"Add<F<0>,Mul<S<0>,Mul<F<1>,F<2>>>>"]
    r, loc = capi.Range(), (C.c_int * 3)()
    assert l.opf_expr_prepare(b"Par<x>", F, 0, capi.R_ACCESSIBLE, C.byref(r), loc) != 0
