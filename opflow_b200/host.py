"""Thin Python mirror of the reference's front-end spellings over the C ABI (used by tests/ and bench.py).

Names follow the reference: MeshBuilder (src/Core/Mesh/Structured/CartesianMesh.hpp:122-304), ExprBuilder /
CartesianField (src/Core/Field/MeshBased/Structured/CartesianField.hpp:38,796-1033), operators d2x/d2y/d2z, dx/dy/dz,
d1IntpCenterToCorner/CornerToCenter (src/Core/Operator/FDMOperators/DiffsInterface.hpp:21-47,
src/Core/Operator/Interpolator/IntpInterface.hpp:26-40).  The product host side is the C++ front-end under
opflow_b200/include; this module only builds signatures/handles and calls libopflow_b200.so -- it never computes
field values itself.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import (BC_ASYMM, BC_DIRC, BC_NEUM, BC_PERIODIC, BC_SYMM, BC_UNDEFINED, LOC_CENTER, LOC_CORNER, MAX_DIM,
                   OP_ADD, OP_DIV, OP_EQ, OP_MINUS, OP_MUL, POS_END, POS_START, Range, check, handle, lib)


class DimPos:
    start, end = POS_START, POS_END


class BCType:
    Undefined, Dirc, Neum, Periodic, Internal, Symm, ASymm = range(7)


class LocOnMesh:
    Corner, Center = LOC_CORNER, LOC_CENTER


# ----------------------------------------------------------------------------------------------- expressions
class Expr:
    """Expression tree node; signature() yields the device-functor type string (include/opflow_b200.h grammar)."""

    def __init__(self, name, children=(), axis=None):
        self.name, self.children, self.axis = name, tuple(children), axis

    # arithmetic overloads: BinOpDefMacros.hpp.in:159-173 (scalar wrapped as ScalarExpr on either side)
    def __add__(self, o): return Expr("Add", (self, _wrap(o)))
    def __radd__(self, o): return Expr("Add", (_wrap(o), self))
    def __sub__(self, o): return Expr("Sub", (self, _wrap(o)))
    def __rsub__(self, o): return Expr("Sub", (_wrap(o), self))
    def __mul__(self, o): return Expr("Mul", (self, _wrap(o)))
    def __rmul__(self, o): return Expr("Mul", (_wrap(o), self))
    def __truediv__(self, o): return Expr("Div", (self, _wrap(o)))
    def __rtruediv__(self, o): return Expr("Div", (_wrap(o), self))
    def __neg__(self): return Expr("Neg", (self,))
    def __gt__(self, o): return Expr("Gt", (self, _wrap(o)))
    def __lt__(self, o): return Expr("Lt", (self, _wrap(o)))
    def __ge__(self, o): return Expr("Ge", (self, _wrap(o)))
    def __le__(self, o): return Expr("Le", (self, _wrap(o)))

    def flatten(self):
        """-> (signature, [field handles in leaf order], [scalars in leaf order]); every leaf occurrence gets a fresh
        index in preorder, exactly like the C++ front-end (a type cannot know that two leaves are the same field)."""
        fields, scalars = [], []

        def rec(e):
            if isinstance(e, Field):
                fields.append(e)
                return f"F<{len(fields) - 1}>"
            if isinstance(e, Scalar):
                scalars.append(float(e.value))
                return f"S<{len(scalars) - 1}>"
            if e.name == "Conv":  # the kernel tensor's entries take consecutive scalar slots, x fastest (FixedSizeTensor order)
                k0 = len(scalars)
                scalars.extend(float(v) for v in e.kernel.reshape(-1, order="F"))
                n = list(e.kernel.shape) + [1] * (3 - e.kernel.ndim)
                return f"Conv<{n[0]},{n[1]},{n[2]},{k0},{rec(e.children[0])}>"
            parts = [rec(c) for c in e.children]
            if e.axis is not None:
                parts.insert(0, str(e.axis))
            return f"{e.name}<{','.join(parts)}>"

        return rec(self), fields, scalars

    def signature(self):
        return self.flatten()[0]


class Scalar(Expr):
    def __init__(self, value):
        super().__init__("S")
        self.value = value


def _wrap(o):
    return o if isinstance(o, Expr) else Scalar(o)


def _stencil(name):
    def make(axis):
        return lambda e: Expr(name, (_wrap(e),), axis=axis)
    return make


class _Kernel:
    """Stand-in for the reference's kernel template-template parameter, e.g. dx(D1FirstOrderCentered, u)."""

    def __init__(self, node):
        self.node = node


D2SecondOrderCentered = _Kernel("D2C")
D1FirstOrderCentered = _Kernel("D1C")
D1FirstOrderBiasedDownwind = _Kernel("D1Dn")
D1FirstOrderBiasedUpwind = _Kernel("D1Up")
D1WENO53Downwind = _Kernel("WenoDn")
D1WENO53Upwind = _Kernel("WenoUp")


def d1(kernel, axis, e): return Expr(kernel.node, (_wrap(e),), axis=axis)
def dx(kernel, e): return d1(kernel, 0, e)
def dy(kernel, e): return d1(kernel, 1, e)
def dz(kernel, e): return d1(kernel, 2, e)
def d2x(kernel, e): return d1(kernel, 0, e)
def d2y(kernel, e): return d1(kernel, 1, e)
def d2z(kernel, e): return d1(kernel, 2, e)
def d1IntpCenterToCorner(axis, e): return Expr("IntpC2N", (_wrap(e),), axis=axis)
def d1IntpCornerToCenter(axis, e): return Expr("IntpN2C", (_wrap(e),), axis=axis)
def conditional(c, a, b): return Expr("Cond", (_wrap(c), _wrap(a), _wrap(b)))
def sqrt(e): return Expr("Sqrt", (_wrap(e),))
def abs_(e): return Expr("Abs", (_wrap(e),))
def pow2(e): return Expr("Pow2", (_wrap(e),))
def max_(a, b): return Expr("Max", (_wrap(a), _wrap(b)))
def min_(a, b): return Expr("Min", (_wrap(a), _wrap(b)))

# the rest of the reference's point-wise family (src/Core/Operator/Arithmetic/AMDS.hpp:40-89)
_UNARY_NODES = ['Exp2', 'Expm1', 'Log10', 'Log2', 'Log1p', 'Cbrt', 'ASin', 'ACos', 'ATan', 'Sinh', 'Cosh', 'ASinh', 'ACosh', 'ATanh', 'Erf', 'Erfc', 'TGamma', 'LGamma', 'Ceil', 'Floor', 'Trunc', 'Round', 'LRound', 'LLRound', 'NearbyInt', 'Rint', 'LRint', 'LLRint', 'ILogb', 'Logb', 'Exp', 'Log', 'Sin', 'Cos', 'Tan', 'Tanh', 'Not', 'Pos']
_BINARY_NODES = ['FMod', 'Remainder', 'FDim', 'Hypot', 'ATan2', 'Ldexp', 'Scalbn', 'Scalbln', 'Nextafter', 'Nexttoward', 'Copysing', 'Pow']


# flux-limiter interpolators (src/Core/Operator/Interpolator/D1FluxLimiterBasedIntpOp.hpp:22-61): d1IntpFl("Quick", axis, "C2N", u, e)
FLUX_LIMITERS = ["Central", "Quick", "Cui", "Fromm", "Lui", "Minmod", "Superbee", "Muscl", "Harmonic", "Albada"]


def d1IntpFl(scheme, axis, direction, u, e):
    """D1Intp<D1QUICK | D1Minmod | ...><axis, Cen2Cor | Cor2Cen>(u, e): face value of e, upwinded by the sign of u"""
    assert scheme in FLUX_LIMITERS and direction in ("C2N", "N2C")
    return Expr(f"Fl{scheme}{direction}", (_wrap(u), _wrap(e)), axis=axis)


def conv(e, kernel):
    """conv(e, kernel) (src/Core/Operator/Convolution/Convolution.hpp): kernel = odd-sized array indexed [i0, i1(, i2)]"""
    k = np.asarray(kernel, dtype=np.float64)
    assert all(n % 2 == 1 for n in k.shape)
    x = Expr("Conv", (_wrap(e),))
    x.kernel = k
    return x


def unary(node, e):
    """point-wise node by name, e.g. unary("Erf", u)"""
    assert node in _UNARY_NODES + ["Sqrt", "Abs", "Pow2", "Neg"], node
    return Expr(node, (_wrap(e),))


def binary(node, a, b):
    assert node in _BINARY_NODES + ["Min", "Max", "Add", "Sub", "Mul", "Div", "Lt", "Le", "Gt", "Ge", "Eq", "Ne", "And", "Or"], node
    return Expr(node, (_wrap(a), _wrap(b)))



# ----------------------------------------------------------------------------------------------- mesh
class CartesianMesh:
    def __init__(self, h, dim):
        self.h, self.dim = h, dim

    def ranges(self):
        r, e = Range(), Range()
        check(lib().opf_mesh_get_range(self.h, C.byref(r), C.byref(e)))
        return r, e

    def axis(self, d):
        """(x, dx, idx) of axis d over the mesh's extended range (CartesianMesh::_x/_dx/_idx)."""
        _, e = self.ranges()
        n = e.end[d] - e.start[d]
        x, dxa, idx = np.zeros(n), np.zeros(n - 1), np.zeros(n - 1)
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        got = lib().opf_mesh_get_axis(self.h, d, p(x), p(dxa), p(idx), n)
        if got != n:
            raise capi.EngineError("opf_mesh_get_axis failed")
        return x, dxa, idx

    def __del__(self):
        try:
            lib().opf_mesh_destroy(self.h)
        except Exception:
            pass


class MeshBuilder:
    """MeshBuilder<CartesianMesh<Meta::int_<dim>>>"""

    def __init__(self, dim):
        self.dim, self.dims, self.start, self.pad = dim, None, [0] * dim, 5
        self.axes, self.ext_modes = {}, {}

    def newMesh(self, *dims):
        assert len(dims) == self.dim
        self.dims = list(dims)
        return self

    def setStart(self, s):
        self.start = list(s)
        return self

    def setPadWidth(self, w):
        self.pad = w
        return self

    def setExtMode(self, mode):
        for d in range(self.dim):
            self.ext_modes[d] = mode
        return self

    def setExtModeOfDim(self, d, mode):
        self.ext_modes[d] = mode
        return self

    def setMeshOfDim(self, k, a, b=None):
        self.axes[k] = (a, b)
        return self

    def build(self):
        I = (C.c_int * self.dim)
        h = handle(lib().opf_mesh_create(self.dim, I(*self.dims), I(*self.start), self.pad), "opf_mesh_create")
        for d, m in self.ext_modes.items():
            check(lib().opf_mesh_set_ext_mode(h, d, m))
        for k, (a, b) in self.axes.items():
            if b is None:  # functor / coordinate array: setMeshOfDim(k, f)
                xs = np.ascontiguousarray([a(i) for i in range(self.start[k], self.start[k] + self.dims[k])] if callable(a)
                                          else a, dtype=np.float64)
                check(lib().opf_mesh_set_coords(h, k, xs.ctypes.data_as(C.POINTER(C.c_double)), len(xs)))
            else:
                check(lib().opf_mesh_set_uniform(h, k, float(a), float(b)))
        return CartesianMesh(h, self.dim)


# ----------------------------------------------------------------------------------------------- field
class Field(Expr):
    """CartesianField<Real, Mesh>: a device-resident field handle."""

    def __init__(self, h, mesh, name):
        super().__init__("F")
        self.h, self.mesh, self.fname, self.dim = h, mesh, name, mesh.dim

    # ranges -----------------------------------------------------------------
    def _range(self, which):
        r = Range()
        check(lib().opf_field_get_range(self.h, which, C.byref(r)))
        return r

    localRange = property(lambda s: s._range(capi.R_LOCAL))
    assignableRange = property(lambda s: s._range(capi.R_ASSIGNABLE))
    accessibleRange = property(lambda s: s._range(capi.R_ACCESSIBLE))
    logicalRange = property(lambda s: s._range(capi.R_LOGICAL))
    storageRange = property(lambda s: s._range(capi.R_STORAGE))

    def getLocalReadableRange(self):
        return self._range(capi.R_READABLE)

    @property
    def loc(self):
        a = (C.c_int * MAX_DIM)()
        check(lib().opf_field_get_loc(self.h, a))
        return list(a)[:self.dim]

    @property
    def padding(self):
        return lib().opf_field_padding(self.h)

    # transfers --------------------------------------------------------------
    def to_numpy(self, r: Range | None = None):
        """Values over `r` (default localRange) as an array indexed [i0, i1, i2] (Fortran order: axis 0 fastest)."""
        r = r or self.localRange
        shape = r.shape(self.dim)
        out = np.empty(shape, dtype=np.float64, order="F")
        check(lib().opf_field_download(self.h, C.byref(r), out.ctypes.data_as(C.c_void_p)))
        return out

    def from_numpy(self, a, r: Range | None = None, update_padding=True):
        r = r or self.localRange
        a = np.asfortranarray(a, dtype=np.float64)
        assert a.shape == r.shape(self.dim), (a.shape, r.shape(self.dim))
        check(lib().opf_field_upload(self.h, C.byref(r), a.ctypes.data_as(C.c_void_p)))
        if update_padding:
            self.updatePadding()
        return self

    def upload_raw(self, ptr, r: Range):
        """host pointer (e.g. pinned torch tensor) -> device, asynchronous on the engine stream"""
        check(lib().opf_field_upload(self.h, C.byref(r), C.c_void_p(ptr)))

    def download_raw(self, ptr, r: Range):
        check(lib().opf_field_download(self.h, C.byref(r), C.c_void_p(ptr)))

    def initBy(self, f):
        """CartesianField::initBy (CartesianField.hpp:283-294): the functor is arbitrary host code, so it is evaluated on
        the host at the reference's coordinates (x or x + .5*dx per loc) and uploaded."""
        a, l = self.assignableRange, self.localRange
        s = [max(a.start[d], l.start[d]) for d in range(self.dim)]
        e = [min(a.end[d], l.end[d]) for d in range(self.dim)]
        _, ext = self.mesh.ranges()
        coords = []
        for d in range(self.dim):
            x, dxa, _ = self.mesh.axis(d)
            idx = np.arange(s[d], e[d]) - ext.start[d]
            coords.append(x[idx] if self.loc[d] == LOC_CORNER else x[idx] + .5 * dxa[idx])
        vals = np.empty([e[d] - s[d] for d in range(self.dim)], order="F")
        it = np.nditer(vals, flags=["multi_index"], op_flags=["writeonly"])
        for v in it:
            v[...] = f([coords[d][it.multi_index[d]] for d in range(self.dim)])
        return self.from_numpy(vals, Range.make(s, e))

    # assignment -------------------------------------------------------------
    def assign(self, e, op=OP_EQ):
        """Expr::operator= / += / -= / *= / /= (Expr.hpp:53-117)"""
        if isinstance(e, (int, float)):
            check(lib().opf_field_assign_scalar(self.h, op, float(e)))
            return self
        sig, fields, scalars = _wrap(e).flatten()
        F = (C.c_void_p * max(1, len(fields)))(*[f.h for f in fields])
        S = (C.c_double * max(1, len(scalars)))(*scalars)
        check(lib().opf_assign(self.h, op, sig.encode(), F, len(fields), S, len(scalars)))
        return self

    def __iadd__(self, e): return self.assign(e, OP_ADD)
    def __isub__(self, e): return self.assign(e, OP_MINUS)
    def __imul__(self, e): return self.assign(e, OP_MUL)
    def __itruediv__(self, e): return self.assign(e, OP_DIV)

    def resplit(self, split_map):
        """CartesianField::resplitWithStrategy (CartesianField.hpp:83-177): split_map = the new strategy's cell-centred blocks, one per rank"""
        arr = (Range * len(split_map))(*split_map)
        check(lib().opf_field_resplit(self.h, arr))
        return self

    def resplit_plan(self, split_map):
        """host half of resplit (works on a plan): -> (send boxes per rank, recv boxes per rank, new localRange)"""
        n = len(split_map)
        arr = (Range * n)(*split_map)
        send, recv, nl = (Range * n)(), (Range * n)(), Range()
        check(lib().opf_field_resplit_plan(self.h, arr, send, recv, C.byref(nl)))
        return list(send), list(recv), nl

    def updatePadding(self):
        check(lib().opf_field_update_padding(self.h))

    def clone(self, name=None):
        return Field(handle(lib().opf_field_clone(self.h, (name or self.fname).encode()), "opf_field_clone"), self.mesh,
                     name or self.fname)

    def neighbors(self):
        n = lib().opf_field_neighbors(self.h, 0, None, None, None, None)
        ranks, codes = (C.c_int * max(1, n))(), (C.c_int * max(1, n))()
        send, recv = (Range * max(1, n))(), (Range * max(1, n))()
        lib().opf_field_neighbors(self.h, n, ranks, send, recv, codes)
        return [(ranks[i], send[i].tup(self.dim), recv[i].tup(self.dim), codes[i]) for i in range(n)]

    def __del__(self):
        try:
            lib().opf_field_destroy(self.h)
        except Exception:
            pass


def prepared(e: Expr, which=capi.R_ACCESSIBLE):
    """ranges / loc of the prepared expression (Expr::prepare())"""
    sig, fields, _ = e.flatten()
    F = (C.c_void_p * max(1, len(fields)))(*[f.h for f in fields])
    r, loc = Range(), (C.c_int * MAX_DIM)()
    check(lib().opf_expr_prepare(sig.encode(), F, len(fields), which, C.byref(r), loc))
    return r, list(loc)


def rangeReduce(e: Expr, rop=capi.RED_SUM, r: Range | None = None):
    """rangeReduce(range, op, [&](auto&& i){ return e.evalAt(i); }) (RangeFor.hpp:87-121) on the device"""
    sig, fields, scalars = _wrap(e).flatten()
    F = (C.c_void_p * max(1, len(fields)))(*[f.h for f in fields])
    S = (C.c_double * max(1, len(scalars)))(*scalars)
    out = C.c_double()
    check(lib().opf_reduce(rop, sig.encode(), F, len(fields), S, len(scalars), C.byref(r) if r is not None else None,
                           C.byref(out)))
    return out.value


class ExprBuilder:
    """ExprBuilder<CartesianField<Real, Mesh>> (CartesianField.hpp:796-1033); like the reference, a builder can be reused
    for several fields (LidDriven2D.cpp:13-26)."""

    def __init__(self):
        self.name, self.mesh = "", None
        self.loc = [LOC_CORNER] * MAX_DIM
        self.bc = [[(BC_UNDEFINED, 0.0, None, None), (BC_UNDEFINED, 0.0, None, None)] for _ in range(MAX_DIM)]
        self.ext = [[0, 0] for _ in range(MAX_DIM)]
        self.padding = 0
        self.split = None  # (n_ranks, rank, [Range])

    def setName(self, n):
        self.name = n
        return self

    def setMesh(self, m):
        self.mesh = m
        return self

    def setLoc(self, loc):
        if isinstance(loc, int):
            self.loc = [loc] * MAX_DIM
        else:
            for d, l in enumerate(loc):
                self.loc[d] = l
        return self

    def setLocOfDim(self, d, l):
        self.loc[d] = l
        return self

    def setBC(self, d, pos, type_, val=0.0, face=None, face_range=None):
        self.bc[d][pos] = (type_, float(val), face, face_range)
        return self

    def setExt(self, *a):
        if len(a) == 1:
            for d in range(MAX_DIM):
                self.ext[d] = [a[0], a[0]]
        else:
            d, pos, w = a
            self.ext[d][pos] = w
        return self

    def setPadding(self, p):
        self.padding = p
        return self

    def setSplitStrategy(self, n_ranks, rank, split_map):
        """split_map: list of cell-centred Range per rank = strategy->getSplitMap(mesh.getRange(), plan)"""
        self.split = (n_ranks, rank, split_map)
        return self

    def build(self, plan_only=False):
        d = capi.FieldDesc()
        d.mesh = self.mesh.h
        dim = self.mesh.dim
        keep = []
        for k in range(dim):
            d.loc[k] = self.loc[k]
            for s in range(2):
                t, v, face, fr = self.bc[k][s]
                d.bc[k][s].type, d.bc[k][s].value = t, v
                if face is not None:
                    fa = np.asfortranarray(face, dtype=np.float64)
                    keep.append(fa)
                    d.bc[k][s].face = fa.ctypes.data_as(C.POINTER(C.c_double))
                    d.bc[k][s].face_range = fr
                d.ext[k][s] = self.ext[k][s]
        d.padding = self.padding
        if self.split:
            n, r, sm = self.split
            arr = (Range * n)(*sm)
            keep.append(arr)
            d.n_ranks, d.rank, d.split_map = n, r, arr
        else:
            d.n_ranks, d.rank, d.split_map = 0, 0, None
        h = handle((lib().opf_field_plan if plan_only else lib().opf_field_create)(C.byref(d), self.name.encode()), "opf_field_create")
        return Field(h, self.mesh, self.name)

    def plan(self):
        """ranges / split / neighbour lists only (opf_field_plan): no device storage, works without a GPU"""
        return self.build(plan_only=True)


def split_even(mesh: CartesianMesh, n_ranks):
    """EvenSplitStrategy<F>::getSplitMap(mesh.getRange(), plan) (EvenSplitStrategy.hpp:57-192)"""
    r, _ = mesh.ranges()
    out = (Range * n_ranks)()
    check(lib().opf_split_even(mesh.dim, C.byref(r), n_ranks, out))
    return list(out)


def split_slab(mesh: CartesianMesh, n_ranks):
    r, _ = mesh.ranges()
    out = (Range * n_ranks)()
    check(lib().opf_split_slab(mesh.dim, C.byref(r), n_ranks, out))
    return list(out)


def set_mode(mode):
    check(lib().opf_set_mode(mode))


def synchronize():
    check(lib().opf_synchronize())


# ----------------------------------------------------------------------------------------------- implicit
class Unknown(Expr):
    """the `e` of the reference's equation lambda `[&](auto&& e) { return lhs(e) == rhs; }`"""

    def __init__(self):
        super().__init__("F")


class StructSolverType:
    NONE, Jacobi, SMG, PFMG, CYCRED, PCG, GMRES, FGMRES, LGMRES, BICGSTAB = range(10)


class EqnSolveHandler:
    """makeEqnSolveHandler(f, target, solver) (HYPREEqnSolveHandler.hpp:44-48) over the matrix-free C ABI.

    f(e) must return (lhs, rhs) -- the two sides of the reference's `lhs == rhs`; lhs linear in e, rhs independent of e."""

    def __init__(self, f, target: Field, type_=StructSolverType.PCG, precond=StructSolverType.NONE, tol=1e-10, maxIter=100,
                 staticMat=False, pinValue=False, numPreRelax=1, numPostRelax=1, precondMaxIter=1, relaxType=1):
        e = Unknown()
        self.target, self.e = target, e
        self.lhs, self.rhs = f(e)
        p = capi.SolverParams(type=type_, precond=precond, tol=tol, max_iter=maxIter, static_mat=int(staticMat), pin_value=int(pinValue),
                              precond_tol=0.0, precond_max_iter=precondMaxIter, num_pre_relax=numPreRelax, num_post_relax=numPostRelax,
                              relax_type=relaxType, print_level=0)
        sig, fields, scalars = self._flatten(self.lhs)
        mask = 0
        for k, f_ in enumerate(fields):
            if f_ is e:
                mask |= 1 << k
        F = (C.c_void_p * max(1, len(fields)))(*[None if f_ is e else f_.h for f_ in fields])
        S = (C.c_double * max(1, len(scalars)))(*scalars)
        self.h = handle(lib().opf_solver_create(target.h, sig.encode(), F, len(fields), S, len(scalars), mask, C.byref(p)), "opf_solver_create")
        self.lhs_sig = sig

    @staticmethod
    def _flatten(expr):
        fields, scalars = [], []

        def rec(x):
            if isinstance(x, (Field, Unknown)):
                fields.append(x)
                return f"F<{len(fields) - 1}>"
            if isinstance(x, Scalar):
                scalars.append(float(x.value))
                return f"S<{len(scalars) - 1}>"
            parts = [rec(c) for c in x.children]
            if x.axis is not None:
                parts.insert(0, str(x.axis))
            return f"{x.name}<{','.join(parts)}>"

        return rec(_wrap(expr)), fields, scalars

    def levels(self):
        return lib().opf_solver_levels(self.h)

    def solve(self):
        """-> EqnSolveState(niter, relerr, abserr)"""
        sig, fields, scalars = self._flatten(self.rhs)
        F = (C.c_void_p * max(1, len(fields)))(*[f_.h for f_ in fields])
        S = (C.c_double * max(1, len(scalars)))(*scalars)
        st = capi.SolveState()
        check(lib().opf_solver_solve(self.h, sig.encode(), F, len(fields), S, len(scalars), C.byref(st)))
        return st

    def export_csr(self, pin_last=False, cap_nnz=None):
        """CSRMatrixGenerator::generate (CSRMatrixGenerator.hpp:55-138) -> (ptr, col, val, rhs) numpy arrays of the assembled system."""
        import numpy as np
        sig, fields, scalars = self._flatten(self.rhs)
        F = (C.c_void_p * max(1, len(fields)))(*[f_.h for f_ in fields])
        S = (C.c_double * max(1, len(scalars)))(*scalars)
        rows = 1
        for n in self.target.assignableRange.shape(self.target.dim):
            rows *= n
        cap = int(cap_nnz) if cap_nnz is not None else rows * 32
        ptr = np.zeros(rows + 1, dtype=np.int32)
        col = np.zeros(cap, dtype=np.int32)
        val = np.zeros(cap, dtype=np.float64)
        rhs = np.zeros(rows, dtype=np.float64)
        nnz = C.c_longlong(0)
        check(lib().opf_solver_export_csr(self.h, sig.encode(), F, len(fields), S, len(scalars), int(pin_last), cap,
                                          ptr.ctypes.data_as(C.POINTER(C.c_int)), col.ctypes.data_as(C.POINTER(C.c_int)),
                                          val.ctypes.data_as(C.POINTER(C.c_double)), rhs.ctypes.data_as(C.POINTER(C.c_double)), C.byref(nnz)))
        return ptr, col[:nnz.value].copy(), val[:nnz.value].copy(), rhs

    def __del__(self):
        try:
            lib().opf_solver_destroy(self.h)
        except Exception:
            pass
