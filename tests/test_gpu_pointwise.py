"""Every point-wise node of the device functor library against the oracle (reference: the functor families of
src/Core/Operator/Arithmetic/AMDS.hpp:34-91, MinMax.hpp, Logical/Compare.hpp, Logical/Boolean.hpp; evaluation rule
BinOpDefMacros.hpp.in:15-17 / UniOpDefMacros.hpp.in:15-17: t1.evalAt(i) op t2.evalAt(i), operand order preserved).

Bit-exact where IEEE fixes the result (rounding family, fmod / remainder / fdim, copysign, nextafter, ldexp / scalbn, comparisons,
sqrt, abs); transcendental functions are compared within 8 ulp of the field scale: CUDA's libm and glibc's are both correctly
rounded to 1-2 ulp but not to the same last bit (tgamma / lgamma: 1e-13 relative, the documented CUDA bound is ~10 ulp)."""
import numpy as np
import pytest

from opflow_b200 import capi, host
from helpers import make_pair, make_pair_on, set_both
from oracle import oracle as O

pytestmark = pytest.mark.gpu

EPS = np.finfo(float).eps
# node -> (domain of the operand(s), tolerance: None = bit-exact, else relative to max |reference|)
UNARY = {
    "Neg": ((-3, 3), None), "Pos": ((-3, 3), None), "Not": ((-1, 1), None), "Abs": ((-3, 3), None), "Sqrt": ((0, 9), None), "Pow2": ((-3, 3), None),
    "Exp": ((-3, 3), 8 * EPS), "Exp2": ((-3, 3), 8 * EPS), "Expm1": ((-1, 1), 8 * EPS), "Log": ((0.1, 9), 8 * EPS), "Log10": ((0.1, 9), 8 * EPS),
    "Log2": ((0.1, 9), 8 * EPS), "Log1p": ((-0.5, 3), 8 * EPS), "Cbrt": ((-8, 8), 8 * EPS), "Sin": ((-6, 6), 8 * EPS), "Cos": ((-6, 6), 8 * EPS),
    "Tan": ((-1.2, 1.2), 8 * EPS), "ASin": ((-0.95, 0.95), 8 * EPS), "ACos": ((-0.95, 0.95), 8 * EPS), "ATan": ((-5, 5), 8 * EPS),
    "Sinh": ((-3, 3), 8 * EPS), "Cosh": ((-3, 3), 8 * EPS), "Tanh": ((-3, 3), 8 * EPS), "ASinh": ((-5, 5), 8 * EPS), "ACosh": ((1.1, 9), 8 * EPS),
    "ATanh": ((-0.9, 0.9), 8 * EPS), "Erf": ((-2, 2), 8 * EPS), "Erfc": ((-2, 2), 8 * EPS), "TGamma": ((0.5, 5), 1e-13), "LGamma": ((0.5, 5), 1e-13),
    "Ceil": ((-5, 5), None), "Floor": ((-5, 5), None), "Trunc": ((-5, 5), None), "Round": ((-5, 5), None), "LRound": ((-5, 5), None),
    "LLRound": ((-5, 5), None), "NearbyInt": ((-5, 5), None), "Rint": ((-5, 5), None), "LRint": ((-5, 5), None), "LLRint": ((-5, 5), None),
    "ILogb": ((0.01, 100), None), "Logb": ((0.01, 100), None),
}
BINARY = {
    "Add": ((-3, 3), (-3, 3), None), "Sub": ((-3, 3), (-3, 3), None), "Mul": ((-3, 3), (-3, 3), None), "Div": ((-3, 3), (0.5, 3), None),
    "Min": ((-3, 3), (-3, 3), None), "Max": ((-3, 3), (-3, 3), None), "Pow": ((0.1, 3), (-2, 2), 16 * EPS),
    "FMod": ((-9, 9), (0.5, 3), None), "Remainder": ((-9, 9), (0.5, 3), None), "FDim": ((-3, 3), (-3, 3), None), "Hypot": ((-3, 3), (-3, 3), 8 * EPS),
    "ATan2": ((-3, 3), (-3, 3), 8 * EPS), "Ldexp": ((-3, 3), (-4, 4), None), "Scalbn": ((-3, 3), (-4, 4), None), "Scalbln": ((-3, 3), (-4, 4), None),
    "Nextafter": ((-3, 3), (-3, 3), None), "Nexttoward": ((-3, 3), (-3, 3), None), "Copysing": ((-3, 3), (-3, 3), None),
    "Lt": ((-1, 1), (-1, 1), None), "Le": ((-1, 1), (-1, 1), None), "Ge": ((-1, 1), (-1, 1), None), "Eq": ((-1, 1), (-1, 1), None),
    "Ne": ((-1, 1), (-1, 1), None), "And": ((-1, 1), (-1, 1), None), "Or": ((-1, 1), (-1, 1), None),
}
DIMS = [70, 23]


def fields(n):
    bc = {(d, s): (capi.BC_NEUM, 0.0) for d in range(2) for s in range(2)}
    u, ou = make_pair(DIMS, [0, 0], [1, 1], bc=bc, name="u")
    out = [(u, ou)]
    for k in range(1, n):
        out.append(make_pair_on(u.mesh, ou.mesh, bc=bc, name=f"f{k}"))
    return out


def sample(rng, dom, shape, quantised=False):
    a = rng.uniform(dom[0], dom[1], shape)
    if quantised:  # comparisons / booleans need ties and zeros to be exercised
        a = np.round(a * 2) / 2
    return np.asfortranarray(a)


def check(got, ref, tol, what):
    if tol is None:
        same = (got == ref) | (np.isnan(got) & np.isnan(ref))
        assert same.all(), f"{what}: {np.count_nonzero(~same)} values not bit-identical, e.g. {got[~same][:3]} vs {ref[~same][:3]}"
    else:
        err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)
        assert err <= tol, f"{what}: relative error {err:.3e} > {tol:.1e}"


@pytest.mark.parametrize("mode", [capi.MODE_EXACT, capi.MODE_FAST])
@pytest.mark.parametrize("node", sorted(UNARY))
def test_unary_node(engine, oracle, mode, node):
    host.set_mode(mode)
    dom, tol = UNARY[node]
    (u, ou), (d, od) = fields(2)
    rng = np.random.default_rng(abs(hash(node)) % 2 ** 31)
    a = sample(rng, dom, u.localRange.shape(2), quantised=node in ("Not",))
    if node in ("Ceil", "Floor", "Trunc", "Round", "LRound", "LLRound", "NearbyInt", "Rint", "LRint", "LLRint"):
        a[::3, ::2] = np.round(a[::3, ::2]) + 0.5  # exact ties: round-half-away vs round-half-even
        a[1::5, 1::3] = np.round(a[1::5, 1::3])
    set_both(u, ou, arr=a)
    e = host.unary(node, u)
    d.assign(e)
    oracle.assign(od, e.signature(), [ou], [])
    r = d.assignableRange
    check(d.to_numpy(r), od.view(r.tup(2)), tol, f"{node} mode={mode}")


@pytest.mark.parametrize("mode", [capi.MODE_EXACT, capi.MODE_FAST])
@pytest.mark.parametrize("node", sorted(BINARY))
def test_binary_node(engine, oracle, mode, node):
    host.set_mode(mode)
    da, db, tol = BINARY[node]
    (u, ou), (v, ov), (d, od) = fields(3)
    rng = np.random.default_rng(abs(hash(node)) % 2 ** 31)
    q = node in ("Lt", "Le", "Ge", "Eq", "Ne", "And", "Or", "Min", "Max", "FDim")
    a, b = sample(rng, da, u.localRange.shape(2), q), sample(rng, db, u.localRange.shape(2), q)
    if node in ("Ldexp", "Scalbn", "Scalbln"):
        b = np.asfortranarray(np.round(b))
    set_both(u, ou, arr=a)
    set_both(v, ov, arr=b)
    e = host.binary(node, u, v)
    d.assign(e)
    oracle.assign(od, e.signature(), [ou, ov], [])
    r = d.assignableRange
    # FAST mode may contract nothing here (single operations), so the bit-exact nodes stay bit-exact in both modes
    check(d.to_numpy(r), od.view(r.tup(2)), tol, f"{node} mode={mode}")
