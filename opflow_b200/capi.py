"""ctypes binding of the C ABI in include/opflow_b200.h (libopflow_b200.so).

This is the only way Python reaches the engine: tests/ and bench.py call the same `extern "C"` entry points the
C++ front-end headers (opflow_b200/include/OpFlow) call.  There is no Python/numpy compute path here -- if the shared
library is missing or no CUDA device is visible, the calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libopflow_b200.so")

MAX_DIM = 3

# enums (numeric values identical to the reference's enums; see include/opflow_b200.h)
LOC_CORNER, LOC_CENTER = 0, 1
POS_START, POS_END = 0, 1
BC_UNDEFINED, BC_DIRC, BC_NEUM, BC_PERIODIC, BC_INTERNAL, BC_SYMM, BC_ASYMM = range(7)
OP_EQ, OP_ADD, OP_MINUS, OP_MUL, OP_DIV = range(5)
MESHEXT_UNDEFINED, MESHEXT_SYMM, MESHEXT_PERIODIC, MESHEXT_UNIFORM = range(4)
MODE_EXACT, MODE_FAST, MODE_STENCIL = 0, 1, 2
RED_SUM, RED_MAX, RED_MIN, RED_ABSMAX, RED_SUMSQ = range(5)
R_LOCAL, R_ASSIGNABLE, R_ACCESSIBLE, R_LOGICAL, R_STORAGE, R_READABLE = range(6)
(SOLVER_NONE, SOLVER_JACOBI, SOLVER_SMG, SOLVER_PFMG, SOLVER_CYCRED, SOLVER_PCG, SOLVER_GMRES, SOLVER_FGMRES,
 SOLVER_LGMRES, SOLVER_BICGSTAB) = range(10)


class Range(C.Structure):
    _fields_ = [("start", C.c_int * MAX_DIM), ("end", C.c_int * MAX_DIM)]

    @classmethod
    def make(cls, start, end):
        r = cls()
        for d in range(MAX_DIM):
            r.start[d] = start[d] if d < len(start) else 0
            r.end[d] = end[d] if d < len(end) else 1
        return r

    def tup(self, dim=MAX_DIM):
        return tuple(self.start[d] for d in range(dim)), tuple(self.end[d] for d in range(dim))

    def shape(self, dim=MAX_DIM):
        return tuple(self.end[d] - self.start[d] for d in range(dim))


class BCDesc(C.Structure):
    _fields_ = [("type", C.c_int), ("value", C.c_double), ("face", C.POINTER(C.c_double)), ("face_range", Range)]


class FieldDesc(C.Structure):
    _fields_ = [("mesh", C.c_void_p), ("loc", C.c_int * MAX_DIM), ("bc", (BCDesc * 2) * MAX_DIM),
                ("ext", (C.c_int * 2) * MAX_DIM), ("padding", C.c_int), ("n_ranks", C.c_int), ("rank", C.c_int),
                ("split_map", C.POINTER(Range))]


class SolverParams(C.Structure):
    _fields_ = [("type", C.c_int), ("precond", C.c_int), ("tol", C.c_double), ("max_iter", C.c_int),
                ("static_mat", C.c_int), ("pin_value", C.c_int), ("precond_tol", C.c_double),
                ("precond_max_iter", C.c_int), ("num_pre_relax", C.c_int), ("num_post_relax", C.c_int),
                ("relax_type", C.c_int), ("print_level", C.c_int), ("k_dim", C.c_int)]


class SolveState(C.Structure):
    _fields_ = [("niter", C.c_int), ("relerr", C.c_double), ("abserr", C.c_double)]


class EngineError(RuntimeError):
    pass


_lib = None

_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)
_R = C.POINTER(Range)
_V = C.c_void_p

# name -> (restype, argtypes); every symbol include/opflow_b200.h declares
SIGNATURES = {
    "opf_init": (C.c_int, [C.c_int]),
    "opf_finalize": (C.c_int, []),
    "opf_last_error": (C.c_char_p, []),
    "opf_version": (C.c_char_p, []),
    "opf_device_count": (C.c_int, []),
    "opf_set_mode": (C.c_int, [C.c_int]),
    "opf_get_mode": (C.c_int, []),
    "opf_synchronize": (C.c_int, []),
    "opf_stream": (_V, []),
    "opf_launch_count": (C.c_longlong, []),
    "opf_last_kernel_name": (C.c_char_p, []),
    "opf_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "opf_get_option": (C.c_int, [C.c_char_p]),
    "opf_timer_begin": (C.c_int, []),
    "opf_timer_end": (C.c_int, [C.POINTER(C.c_float)]),
    "opf_mesh_create": (_V, [C.c_int, _I, _I, C.c_int]),
    "opf_mesh_set_ext_mode": (C.c_int, [_V, C.c_int, C.c_int]),
    "opf_mesh_set_uniform": (C.c_int, [_V, C.c_int, C.c_double, C.c_double]),
    "opf_mesh_set_coords": (C.c_int, [_V, C.c_int, _D, C.c_int]),
    "opf_mesh_get_range": (C.c_int, [_V, _R, _R]),
    "opf_mesh_get_axis": (C.c_int, [_V, C.c_int, _D, _D, _D, C.c_int]),
    "opf_mesh_destroy": (C.c_int, [_V]),
    "opf_field_create": (_V, [C.POINTER(FieldDesc), C.c_char_p]),
    "opf_field_plan": (_V, [C.POINTER(FieldDesc), C.c_char_p]),
    "opf_field_clone": (_V, [_V, C.c_char_p]),
    "opf_field_destroy": (C.c_int, [_V]),
    "opf_field_dim": (C.c_int, [_V]),
    "opf_field_get_range": (C.c_int, [_V, C.c_int, _R]),
    "opf_field_get_loc": (C.c_int, [_V, _I]),
    "opf_field_padding": (C.c_int, [_V]),
    "opf_field_device_ptr": (C.c_int, [_V, C.POINTER(_D), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "opf_field_upload": (C.c_int, [_V, _R, _V]),
    "opf_field_download": (C.c_int, [_V, _R, _V]),
    "opf_host_alloc": (_V, [C.c_ulonglong]),
    "opf_host_free": (C.c_int, [_V]),
    "opf_field_snapshot": (_V, [_V, _R, _V]),
    "opf_snapshot_wait": (C.c_int, [_V]),
    "opf_field_assign_scalar": (C.c_int, [_V, C.c_int, C.c_double]),
    "opf_field_assign_field": (C.c_int, [_V, C.c_int, _V]),
    "opf_field_update_padding": (C.c_int, [_V]),
    "opf_field_set_bc_value": (C.c_int, [_V, C.c_int, C.c_int, C.c_double]),
    "opf_field_swap": (C.c_int, [_V, _V]),
    "opf_field_neighbors": (C.c_int, [_V, C.c_int, _I, _R, _R, _I]),
    "opf_expr_register": (C.c_int, [C.c_char_p, _V]),
    "opf_expr_register_abi": (C.c_int, [C.c_char_p, _V, C.c_ulonglong]),
    "opf_expr_is_registered": (C.c_int, [C.c_char_p]),
    "opf_expr_builtin_count": (C.c_int, []),
    "opf_expr_builtin_name": (C.c_char_p, [C.c_int]),
    "opf_expr_prepare": (C.c_int, [C.c_char_p, C.POINTER(_V), C.c_int, C.c_int, _R, _I]),
    "opf_field_resplit": (C.c_int, [_V, C.POINTER(Range)]),
    "opf_field_resplit_plan": (C.c_int, [_V, C.POINTER(Range), C.POINTER(Range), C.POINTER(Range), C.POINTER(Range)]),
    "opf_assign": (C.c_int, [_V, C.c_int, C.c_char_p, C.POINTER(_V), C.c_int, _D, C.c_int]),
    "opf_assign_ex": (C.c_int, [_V, C.c_int, C.c_char_p, C.POINTER(_V), C.c_int, _D, C.c_int, C.c_int]),
    "opf_assign_repeat": (C.c_int, [_V, C.c_int, C.c_char_p, C.POINTER(_V), C.c_int, _D, C.c_int, C.c_int]),
    "opf_assign_host": (C.c_int, [_V, C.c_int, C.c_char_p, C.POINTER(_V), C.c_int, _D, C.c_int, _V, C.c_void_p, C.c_void_p]),
    "opf_solver_create": (_V, [_V, C.c_char_p, C.POINTER(_V), C.c_int, _D, C.c_int, C.c_uint, C.POINTER(SolverParams)]),
    "opf_solver_solve": (C.c_int, [_V, C.c_char_p, C.POINTER(_V), C.c_int, _D, C.c_int, C.POINTER(SolveState)]),
    "opf_solver_update": (C.c_int, [_V, C.POINTER(_V), C.c_int, _D, C.c_int]),
    "opf_solver_levels": (C.c_int, [_V]),
    "opf_solver_export_csr": (C.c_int, [_V, C.c_char_p, C.POINTER(_V), C.c_int, _D, C.c_int, C.c_int, C.c_longlong, _I, _I, _D, _D, C.POINTER(C.c_longlong)]),
    "opf_solver_destroy": (C.c_int, [_V]),
    "opf_reduce": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(_V), C.c_int, _D, C.c_int, _R, _D]),
    "opf_split_even": (C.c_int, [C.c_int, _R, C.c_int, _R]),
    "opf_split_slab": (C.c_int, [C.c_int, _R, C.c_int, _R]),
    "opf_comm_unique_id": (C.c_int, [_V]),
    "opf_comm_init": (C.c_int, [C.c_int, C.c_int, _V]),
    "opf_comm_rank": (C.c_int, []),
    "opf_comm_size": (C.c_int, []),
    "opf_comm_allreduce": (C.c_int, [_D, C.c_int, C.c_int]),
    "opf_comm_finalize": (C.c_int, []),
}


def lib():
    """Load libopflow_b200.so (built by `make -C opflow_b200` / __graft_entry__.build()).  Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(opflow_b200 has no Python/CPU fallback)")
        l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().opf_last_error()
        raise EngineError(f"opflow_b200 error {rc}: {msg.decode() if msg else ''}")
    return rc


def handle(p, what):
    if not p:
        msg = lib().opf_last_error()
        raise EngineError(f"{what} failed: {msg.decode() if msg else ''}")
    return p
