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
