"""cfem_b200 — B200-native residual-viscosity (RV) P1 hot path.

Host-side mirror of the reference's ``Code/Utils`` entry points and time loops;
all arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI in
``include/cfem_b200.h``.  No CPU fallback.
"""
from . import meshes  # noqa: F401
from ._lib import CfemError, LIB_PATH  # noqa: F401
from .context import Context, step_params  # noqa: F401
from .solvers import (NodalFunction, solve_advection, solve_burgers, solve_kpp,  # noqa: F401
                      kpp_initial_condition, burgers_initial_condition, advection_initial_condition,
                      advection_velocity, advection_dt, solve_euler, sod_initial_condition, l2_error, convergence_rate)
