"""ctypes binding of ``libcfem_b200.so`` (C ABI declared in ``include/cfem_b200.h``).

There is no CPU fallback: if the shared library has not been built (see
``__graft_entry__.build()`` / ``csrc/Makefile``) importing this module raises,
and every compute call raises ``CfemError`` when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcfem_b200.so")

# enums (mirror include/cfem_b200.h)
FLUX_ADVECTION, FLUX_BURGERS, FLUX_KPP = 0, 1, 2
BDF1, BDF2 = 1, 2
EPS_NONLINEAR, EPS_LINEAR, EPS_POINTWISE, EPS_FIRST_ORDER, EPS_LINEAR_SIMPLE, EPS_CELL = 0, 1, 2, 3, 4, 5
MAT_MASS, MAT_MASS_BC, MAT_SYSTEM, MAT_STIFFNESS = 0, 1, 2, 3
SOLVER_PCG, SOLVER_BICGSTAB, SOLVER_GMRES, SOLVER_CHEBYSHEV = 0, 1, 2, 3
BC_CONSTANT, BC_BURGERS_EXACT, BC_USER = 0, 1, 2
ORDER_HILBERT, ORDER_NATURAL = 0, 1
KERNEL_SPMV, KERNEL_ASM_RESIDUAL, KERNEL_ASM_JACOBIAN, KERNEL_RV_EPSILON, KERNEL_ASM_RV_RHS = 0, 1, 2, 3, 4
KERNEL_COMM_ALLREDUCE, KERNEL_COMM_HALO, KERNEL_SPMV_SYSTEM, KERNEL_CHEB_ITER, KERNEL_KRYLOV_ITER = 6, 7, 8, 9, 10

FLUX_BY_NAME = {"advection": FLUX_ADVECTION, "burgers": FLUX_BURGERS, "kpp": FLUX_KPP}
SOLVER_BY_NAME = {"pcg": SOLVER_PCG, "bicgstab": SOLVER_BICGSTAB, "gmres": SOLVER_GMRES, "chebyshev": SOLVER_CHEBYSHEV}


class CfemError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cfem_b200 error {code}: {msg}")
        self.code = code


class StepParams(C.Structure):
    _fields_ = [
        ("flux", C.c_int32), ("scheme", C.c_int32), ("dt", C.c_double),
        ("Cvel", C.c_double), ("Crv", C.c_double),
        ("newton_rtol", C.c_double), ("newton_atol", C.c_double),
        ("newton_max_it", C.c_int32), ("solver", C.c_int32),
        ("lin_rtol", C.c_double), ("lin_max_it", C.c_int32), ("bc_kind", C.c_int32),
        ("bc_value", C.c_double), ("residual_bc", C.c_int32), ("mass_solver", C.c_int32),
        ("mass_rtol", C.c_double),
    ]


class StepStats(C.Structure):
    _fields_ = [
        ("steps", C.c_int64), ("newton_iterations", C.c_int64), ("mass_iterations", C.c_int64),
        ("krylov_iterations", C.c_int64), ("spmv_launches", C.c_int64),
        ("assembly_launches", C.c_int64), ("kernel_launches", C.c_int64),
        ("last_newton_residual", C.c_double), ("time", C.c_double), ("device_ms", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_P = C.c_void_p
_I = C.c_int
_D = C.c_double
_L = C.c_int64

# name -> (restype, argtypes); every symbol include/cfem_b200.h declares
SIGNATURES = {
    "cfem_last_error": (C.c_char_p, []),
    "cfem_version": (_I, []),
    "cfem_struct_size": (_I, [_I]),
    "cfem_device_count": (_I, []),
    "cfem_create": (_I, [C.POINTER(_P), _I, _L, _L, _P, _I, _P, _I, _I]),
    "cfem_nccl_unique_id": (_I, [_P]),
    "cfem_create_distributed": (_I, [C.POINTER(_P), _I, _I, _I, _P, _L, _L, _P, _I, _P, _I, _I]),
    "cfem_num_owned": (_L, [_P]),
    "cfem_num_ghosts": (_L, [_P]),
    "cfem_comm_stats": (_I, [_P, C.POINTER(_L), C.POINTER(_L), C.POINTER(_L)]),
    "cfem_destroy": (None, [_P]),
    "cfem_synchronize": (_I, [_P]),
    "cfem_num_nodes": (_L, [_P]),
    "cfem_num_cells": (_L, [_P]),
    "cfem_num_nonzeros": (_L, [_P]),
    "cfem_num_boundary": (_L, [_P]),
    "cfem_num_dirichlet": (_L, [_P]),
    "cfem_num_tiles": (_L, [_P]),
    "cfem_device_bytes": (_L, [_P]),
    "cfem_device_limits": (_I, [_I, _P]),
    "cfem_comm_timers": (_I, [_P, _P, _I]),
    "cfem_l2_error_p3": (_I, [_P, _P, _P, _P]),
    "cfem_get_csr_pattern": (_I, [_P, _P, _P]),
    "cfem_get_boundary_dofs": (_I, [_P, _P]),
    "cfem_set_dirichlet": (_I, [_P, _P, _L]),
    "cfem_get_ordering": (_I, [_P, _P]),
    "cfem_nodal_h": (_I, [_P, _P, _D, _I, C.POINTER(_I)]),
    "cfem_rv_residual": (_I, [_P, _I, _I, _D, _P, _P, _P, _P, _I, _P, _D, _I, C.POINTER(_I)]),
    "cfem_rv_epsilon": (_I, [_P, _I, _I, _D, _D, _P, _P, _P, _P, _P, _P]),
    "cfem_si_epsilon": (_I, [_P, _I, _D, _D, _I, _P, _P, _P, _P, _P]),
    "cfem_assemble_advection": (_I, [_P, _D, _P, _P, _P, _P, _P]),
    "cfem_assemble_cn_residual": (_I, [_P, _I, _D, _P, _P, _P, _P, _P]),
    "cfem_assemble_cn_jacobian": (_I, [_P, _I, _D, _P, _P]),
    "cfem_assemble_stiffness": (_I, [_P, _P]),
    "cfem_matrix_values": (_I, [_P, _I, _P]),
    "cfem_spmv": (_I, [_P, _I, _P, _P]),
    "cfem_solve": (_I, [_P, _I, _I, _P, _P, _D, _D, _I, C.POINTER(_I), C.POINTER(_D)]),
    "cfem_state_set": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _D]),
    "cfem_state_update": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _D]),
    "cfem_state_get": (_I, [_P, _P, _P, _P, _P, _P, _P, C.POINTER(_D)]),
    "cfem_state_update_owned": (_I, [_P, _P, _P, _P, _P, _P, _D]),
    "cfem_state_get_owned": (_I, [_P, _P, _P, _P, _P, _P, _P, C.POINTER(_D)]),
    "cfem_step_scalar": (_I, [_P, C.POINTER(StepParams), _I, _P, C.POINTER(StepStats)]),
    "cfem_step_advection": (_I, [_P, C.POINTER(StepParams), _I, _I, C.POINTER(StepStats)]),
    "cfem_step_scalar_si": (_I, [_P, C.POINTER(StepParams), _D, _D, _D, _P, _I, _P, C.POINTER(StepStats)]),
    "cfem_smooth_vector": (_I, [_P, _P, _P, _D]),
    "cfem_profile_gaps": (_I, [_P, C.POINTER(_D)]),
    "cfem_euler_state_set": (_I, [_P, _P, _P, _P, _P, _P, _P, _D]),
    "cfem_euler_state_get": (_I, [_P, _P, _P, _P, C.POINTER(_D)]),
    "cfem_step_euler": (_I, [_P, C.POINTER(StepParams), _I, C.POINTER(StepStats)]),
    "cfem_profile_begin": (_I, [_P, _I]),
    "cfem_profile_end": (_I, [_P, C.POINTER(_D), C.POINTER(_L)]),
    "cfem_time_kernel": (_I, [_P, _I, _I, _I, C.POINTER(_D), C.POINTER(_D)]),
    "cfem_host_analyse": (_I, [C.POINTER(_P), _L, _L, _P, _I, _P, _I, _I]),
    "cfem_host_analyse_part": (_I, [C.POINTER(_P), _I, _I, _L, _L, _P, _I, _P, _I, _I]),
    "cfem_host_info": (_L, [_P, _I]),
    "cfem_host_size": (_L, [_P, _I]),
    "cfem_host_copy": (_I, [_P, _I, _P]),
    "cfem_host_analyse_partitioned": (_I, [_P, _I, _I, _L, _L, _P, _I, _P, _I, _I, _P]),
    "cfem_host_partition": (_I, [_I, _I, _L, _L, _P, _I, _P, _I, _P]),
    "cfem_create_partitioned": (_I, [_P, _I, _I, _I, _P, _L, _L, _P, _I, _P, _I, _I, _P]),
    "cfem_host_free": (None, [_P]),
}

HM_ARRAYS = {"n2u": 0, "cells": 1, "rowptr": 2, "colidx": 3, "v2c_ptr": 4, "v2c_code": 5, "tile_node": 6,
             "tile_cellptr": 7, "tile_cells": 8, "is_bnd": 9, "bnd_user": 10, "peer_rank": 11, "send_ptr": 12,
             "send_idx": 13, "recv_off": 14, "recv_cnt": 15, "last_cell": 16, "lc16": 17, "tile_extptr": 18,
             "tile_ext": 19, "tile_order": 20}


def device_limits(device=0):
    """{'l2_bytes', 'persisting_l2_max', 'access_window_max', 'sm_count'} of a CUDA device."""
    out = (C.c_int64 * 4)()
    check(load().cfem_device_limits(int(device), out))
    return dict(zip(("l2_bytes", "persisting_l2_max", "access_window_max", "sm_count"), [int(v) for v in out]))


PART_HILBERT, PART_METIS = 0, 1


def host_partition(x, cells, world, method="metis"):
    """One part id per node (caller numbering): METIS k-way on the nodal graph, or equal Hilbert ranges."""
    lib = load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    cells = np.ascontiguousarray(cells)
    ib = 8 if cells.dtype == np.int64 else 4
    if ib == 4:
        cells = np.ascontiguousarray(cells, dtype=np.int32)
    part = np.empty(x.shape[0], dtype=np.int32)
    check(lib.cfem_host_partition(PART_METIS if method == "metis" else PART_HILBERT, int(world), x.shape[0],
                                  cells.shape[0], ptr(x), x.shape[1], ptr(cells), ib, ptr(part)))
    return part


def host_analyse(x, cells, order=ORDER_HILBERT, rank=0, world=1, node_part=None):
    """Run the once-per-mesh host analysis (no GPU needed) and return its arrays.

    With ``world > 1``: rank's part of the partition (local numbering: owned nodes, then ghosts);
    ``node_part``: the caller's partition (``host_partition``), default equal Hilbert ranges."""
    lib = load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    cells = np.ascontiguousarray(cells)
    ib = 8 if cells.dtype == np.int64 else 4
    if ib == 4:
        cells = np.ascontiguousarray(cells, dtype=np.int32)
    h = C.c_void_p()
    node_part = None if node_part is None else np.ascontiguousarray(node_part, dtype=np.int32)
    check(lib.cfem_host_analyse_partitioned(C.byref(h), rank, world, x.shape[0], cells.shape[0], ptr(x), x.shape[1],
                                            ptr(cells), ib, order, ptr(node_part)))
    out = {}
    try:
        for name, what in HM_ARRAYS.items():
            n = lib.cfem_host_size(h, what)
            dt = {"is_bnd": np.uint8, "v2c_code": np.uint32, "lc16": np.uint16}.get(name, np.int32)
            a = np.empty(n, dtype=dt)
            check(lib.cfem_host_copy(h, what, ptr(a)))
            out[name] = a
        for k, name in enumerate(("n_owned", "n_local", "n_global", "n_cells", "nnz")):
            out[name] = int(lib.cfem_host_info(h, k))
    finally:
        lib.cfem_host_free(h)
    return out

_lib = None


def load():
    """Load the shared library once; raise if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("CFEM_LIB", LIB_PATH)   # CFEM_LIB: an alternative build of the same ABI (kernel A/B runs)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C conservation-fem_b200/csrc`). cfem_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise CfemError(code, load().cfem_last_error().decode("utf-8", "replace"))


def ptr(a):
    """Raw address of a numpy array / torch tensor (host or CUDA) or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch.Tensor
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return a.data_ptr()
    raise TypeError(f"cannot take the address of {type(a)}")


def f64(a):
    """View/copy of ``a`` as a contiguous float64 numpy array (torch tensors pass through)."""
    if a is None:
        return None
    if hasattr(a, "data_ptr") and not isinstance(a, np.ndarray):
        import torch

        if a.dtype != torch.float64:
            raise TypeError("tensors must be float64")
        return a.contiguous()
    return np.ascontiguousarray(a, dtype=np.float64)
