"""GPU: the CUDA path against dolfinx's own output.

The reference keeps three 285-frame dolfinx time series of the linear-advection RV / SI solvers on
its 1,011-node unit-disk mesh (``Code/Linear_advection/Data/RV/RV_node.h5``, ``RV_cell.h5``,
``Data/SI/smoothness.h5``; 13 frames of each are committed in ``tests/golden/ref_series_*.npz``, see
``make_golden.py``).  Here every step of those runs is redone through the C ABI -- residual
projection, viscosity, Crank-Nicolson assembly with lifting, Krylov solve -- and the fields are
compared with the stored frames.  No oracle in this file: it is CUDA vs dolfinx.

Tolerance 1e-10 relative L2 (north-star bar) on every stored frame, up to 285 steps in.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from cfem_b200 import Context, _lib as L  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-10
CVEL, CRV = 0.25, 1.0   # tests/eps_func.py:78-79, RV_cell.py:78-79


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


@pytest.fixture(scope="module")
def disk():
    d = np.load(os.path.join(GOLD, "rv_node_mesh.npz"))
    x, c = d["x"], d["cells"]
    ctx = Context((x, c))
    w = np.stack([-2 * np.pi * x[:, 1], 2 * np.pi * x[:, 0]], axis=1)   # eps_func.py:47-48
    dt = 0.5 * (1 / 16) / np.abs(w).sum(axis=1).max()                   # eps_func.py:69-74 (matrix inf-norm)
    assert dt == float(d["first_time_stamp"])
    u0 = ((x[:, 0] - 0.3) ** 2 + x[:, 1] ** 2 <= 0.25 ** 2).astype(np.float64)   # eps_func.py:44-45
    yield ctx, x, c, w, dt, u0
    ctx.close()


def cn_step(ctx, dt, w, eps, u_n):
    """assemble (matrix + lifted rhs, eps_func.py:205-224) and solve (``:227``)."""
    b = ctx.assemble_advection(dt, w, eps, u_n)
    return ctx.solve(L.MAT_SYSTEM, b, x0=u_n, solver="bicgstab", rtol=1e-14, max_it=4000)


def check(frames, g, name):
    worst = 0.0
    for k, F in zip(g["index"], g["frames"]):
        e = rel(frames[k], F)
        worst = max(worst, e)
        assert e < TOL, (name, int(k), e)
    return worst


@pytest.mark.parametrize("variant", ["eps_func", "rv_cell"])
def test_rv_series_against_dolfinx(disk, variant):
    ctx, x, c, w, dt, u0 = disk
    g = np.load(os.path.join(GOLD, f"ref_series_{variant}.npz"))
    h = ctx.nodal_h()
    u_old = u0.copy()
    u_n = cn_step(ctx, dt, w, None, u0)        # one GFEM step
    frames = [u_n.copy()]
    for _ in range(284):
        Rh = ctx.rv_residual("advection", "bdf1", dt, u_n, u_old, w=w, use_bc=False)
        if variant == "eps_func":
            # Rh / max(u_n - mean) then min(Cvel h |w|, Crv h^2 |R|): get_epsilon_linear_simple divides by
            # ||u_n - mean||_inf, the same number while the pulse is mostly positive
            assert (u_n - u_n.mean()).max() == np.abs(u_n - u_n.mean()).max()
            eps = ctx.rv_epsilon("linear_simple", "advection", CVEL, CRV, u_n=u_n, Rh=Rh, h=h, w=w)
        else:
            eps = ctx.rv_epsilon("cell", "advection", CVEL, CRV, u_n=u_n, Rh=Rh, w=w)
        uh = cn_step(ctx, dt, w, eps, u_n)
        u_old, u_n = u_n, uh
        frames.append(uh.copy())
    check(frames, g, variant)


def test_si_series_against_dolfinx(disk):
    """``smoothness_old_convergence.py:184-253`` as it ran: alpha-weights read from the matrix bound to the
    name ``A`` -- the bc'd unit stiffness matrix for the first SI step, the previous step's assembled
    system matrix afterwards.  Both matrices come from the GPU (``cfem_matrix_values``); only the
    five-line alpha ratio is numpy."""
    ctx, x, c, w, dt, u0 = disk
    g = np.load(os.path.join(GOLD, "ref_series_si_old.npz"))
    h = ctx.nodal_h()
    wn = np.sqrt(w[:, 0] ** 2 + w[:, 1] ** 2)
    bnd = ctx.boundary_dofs()

    def alpha(Kw, u):
        Kw = Kw.tocsr()
        rows = np.repeat(np.arange(Kw.shape[0]), np.diff(Kw.indptr))
        du = u[Kw.indices] - u[rows]
        num = np.bincount(rows, Kw.data * du, Kw.shape[0])
        den = np.bincount(rows, np.abs(Kw.data) * np.abs(du), Kw.shape[0])
        return np.abs(num) / np.maximum(den, 1e-8)

    ctx.assemble_stiffness()
    K = ctx.matrix(L.MAT_STIFFNESS).tolil()
    K[bnd, :] = 0.0
    K[:, bnd] = 0.0
    K[bnd, bnd] = 1.0
    Kw = K.tocsr()
    u_n = cn_step(ctx, dt, w, None, u0)
    frames = [u_n.copy()]
    for _ in range(284):
        eps = alpha(Kw, u_n) * 0.05 * h * wn
        u_n = cn_step(ctx, dt, w, eps, u_n)
        Kw = ctx.matrix(L.MAT_SYSTEM)
        frames.append(u_n.copy())
    check(frames, g, "si_old")
