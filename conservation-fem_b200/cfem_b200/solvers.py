"""Solver entry points: the reference's time loops with the per-step work on the GPU.

Each function mirrors one reference script and keeps taking a dolfinx mesh (or an
``(x, cells)`` pair) and an initial condition (callable on ``x`` with shape
``(3|2, N)`` like ``fem.Function.interpolate``, or a nodal array):

* ``solve_kpp``        ``Code/KPP/KPP_exact.py:47-166``
* ``solve_burgers``    ``Code/Burgers_equation/Exact_Burger_RV.py:28-237``
* ``solve_advection``  ``Code/Linear_advection/RV_node_convergence.py:48-236``
                       (``residual_bc=True``: ``RV_node.py:213``)

``dt`` / ``num_steps`` are explicit arguments because the reference derives them
from round-off-sensitive quantities.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .context import Context, step_params


class _X:
    def __init__(self, array):
        self.array = array


class NodalFunction:
    """Minimal stand-in for ``dolfinx.fem.Function`` on P1: ``.x.array``, ``.name``."""

    def __init__(self, array, name="f"):
        self.x = _X(np.asarray(array, dtype=np.float64))
        self.name = name

    def __array__(self, dtype=None, copy=None):
        return self.x.array if dtype is None else self.x.array.astype(dtype)


def _interpolate(ctx: Context, f):
    """``Function.interpolate``: callables get coordinates as a (3, N) array."""
    if callable(f):
        X = np.zeros((3, ctx.n))
        X[0], X[1] = ctx.x[:, 0], ctx.x[:, 1]
        v = np.asarray(f(X))
        if v.ndim == 2:  # vector-valued (2, N) -> interleaved (N, 2)
            return np.ascontiguousarray(v[:2].T, dtype=np.float64)
        return np.ascontiguousarray(v, dtype=np.float64)
    if hasattr(f, "x") and hasattr(f.x, "array"):
        f = f.x.array
    return np.ascontiguousarray(f, dtype=np.float64)


# ---- problem data of the reference scripts ------------------------------------
def kpp_initial_condition(x):
    """``Code/KPP/KPP_exact.py:52-53``."""
    return (x[0] ** 2 + x[1] ** 2 <= 1) * 14 * np.pi / 4 + (x[0] ** 2 + x[1] ** 2 > 1) * np.pi / 4


def burgers_initial_condition(x):
    """``Code/Burgers_equation/Exact_Burger_RV.py:70-80``."""
    x0, x1 = x[0], x[1]
    u = np.zeros_like(x0)
    u = np.where((x0 <= 0.5) & (x1 >= 0.5), -0.2, u)
    u = np.where((x0 > 0.5) & (x1 >= 0.5), -1.0, u)
    u = np.where((x0 <= 0.5) & (x1 < 0.5), 0.5, u)
    u = np.where((x0 > 0.5) & (x1 < 0.5), 0.8, u)
    return u


def burgers_exact_solution(x, t=0.5):
    """Exact solution of the 2-D Burgers Riemann problem, ``Code/Burgers_equation/Exact_Burger_RV.py:37-66`` /
    ``Exact_Burger_RV_conv.py:31-60``; ``x`` has shape ``(3|2, N)`` like ``Function.interpolate`` passes it.
    (At ``t = 0`` the fan branch divides by zero exactly as the reference's expression does; the NaN comparisons
    are False and the other branches give the initial data.)"""
    X, Y = x[0], x[1]
    u = np.zeros_like(X)
    with np.errstate(all="ignore"):
        m1 = X <= (1 / 2 - 3 * t / 5)
        u = np.where(m1 & (Y > (1 / 2 + 3 * t / 20)), -0.2, u)
        u = np.where(m1 & (Y <= (1 / 2 + 3 * t / 20)), 0.5, u)
        m2 = ((1 / 2 - 3 * t / 5) <= X) & (X <= (1 / 2 - t / 4))
        l2 = -8 * X / 7 + 15 / 14 - 15 * t / 28
        u = np.where(m2 & (Y > l2), -1, u)
        u = np.where(m2 & (Y <= l2), 0.5, u)
        m3 = (1 / 2 - t / 4 <= X) & (X <= (1 / 2 + t / 2))
        l3 = X / 6 + 5 / 12 - 5 * t / 24
        u = np.where(m3 & (Y > l3), -1, u)
        u = np.where(m3 & (Y <= l3), 0.5, u)
        m4 = (1 / 2 + t / 2 <= X) & (X <= (1 / 2 + 4 * t / 5))
        l4 = X - 5 / (18 * t) * (X + t - 1 / 2) ** 2 if t != 0 else X - np.float64(5) / np.float64(0.0) * (X + t - 1 / 2) ** 2
        u = np.where(m4 & (Y > l4), -1, u)
        u = np.where(m4 & (Y <= l4), (2 * X - 1) / (2 * t) if t != 0 else (2 * X - 1) / np.float64(0.0), u)
        m5 = X >= (1 / 2 + 4 * t / 5)
        u = np.where(m5 & (Y > (1 / 2 - t / 10)), -1, u)
        u = np.where(m5 & (Y <= (1 / 2 - t / 10)), 0.8, u)
    return u


def advection_initial_condition(x, r0=0.25, x0_1=0.3, x0_2=0):
    """``Code/Linear_advection/RV_node.py:54-55``."""
    return 1 / 2 * (1 - np.tanh(((x[0] - x0_1) ** 2 + (x[1] - x0_2) ** 2) / r0 ** 2 - 1))


def advection_velocity(x):
    """``Code/Linear_advection/RV_node.py:59-60``."""
    return np.array([-2 * np.pi * x[1], 2 * np.pi * x[0]])


def advection_dt(w, hmax, CFL=0.5):
    """``RV_node.py:78-85`` (matrix inf-norm of the (N,2) velocity table)."""
    return CFL * hmax / np.linalg.norm(np.asarray(w).reshape(-1, 2), ord=np.inf)


# ---- loops ------------------------------------------------------------------------
def _step_user_bc(ctx, p, num_steps, bc_of_step):
    """``num_steps`` steps with caller-supplied Dirichlet values: ``bc_of_step(k)`` -> values on ``ctx.boundary_dofs()``
    for step k (0-based).  The library stages at most 2 N values per call, so the steps go in chunks."""
    nb = ctx.boundary_dofs().size
    chunk = max(1, (2 * ctx.n) // max(nb, 1))
    stats, done = None, 0
    while done < num_steps:
        k = min(chunk, num_steps - done)
        vals = np.ascontiguousarray(np.stack([bc_of_step(done + j) for j in range(k)]), dtype=np.float64)
        st = ctx.step_scalar(p, k, bc_values=vals)
        done += k
        if stats is None:
            stats = st
        else:
            for key in ("steps", "newton_iterations", "krylov_iterations", "mass_iterations", "kernel_launches",
                        "spmv_launches", "assembly_launches", "device_ms"):
                stats[key] += st[key]
            stats["time"], stats["last_newton_residual"] = st["time"], st.get("last_newton_residual")
    return stats


def _run_scalar(flux, domain, u0, dt, num_steps, Cvel, Crv, bc_kind, bc_value, scheme, newton_rtol,
                solver, lin_rtol, device, h, return_stats, xdmf=None, write_every=1, mass_rtol=0.0, bc_of_step=None):
    ctx = domain if isinstance(domain, Context) else Context.for_domain(domain, device=device)
    u0 = _interpolate(ctx, u0)
    h = ctx.nodal_h() if h is None else _interpolate(ctx, h)
    ctx.state_set(uh=u0, u_n=u0, u_old=u0, u_oo=u0, RH=np.zeros(ctx.n), h=h, t=0.0)
    p = step_params(flux, dt, Cvel, Crv, scheme=scheme, newton_rtol=newton_rtol, solver=solver,
                    lin_rtol=lin_rtol, bc_kind=bc_kind, bc_value=bc_value, mass_rtol=mass_rtol)
    if bc_of_step is not None:
        stats = _step_user_bc(ctx, p, num_steps, bc_of_step)
    elif xdmf is None:
        stats = ctx.step_scalar(p, num_steps)
    else:
        # xdmf.write_mesh(domain); xdmf.write_function(uh, t) every `write_every` steps (KPP_exact.py:108-109,165)
        from .io import XdmfWriter

        writer = xdmf if isinstance(xdmf, XdmfWriter) else XdmfWriter(xdmf, ctx.x, ctx.cells)
        stats, done = None, 0
        while done < num_steps:
            k = min(int(write_every), num_steps - done)
            st = ctx.step_scalar(p, k)
            done += k
            writer.write_function(ctx.state_get(("uh",))["uh"], st["time"], name="uh")
            if stats is None:
                stats = st
            else:
                for key in ("steps", "newton_iterations", "krylov_iterations", "mass_iterations", "kernel_launches",
                            "spmv_launches", "assembly_launches", "device_ms"):
                    if key in stats:
                        stats[key] += st[key]
                stats["time"] = st["time"]
                stats["last_newton_residual"] = st.get("last_newton_residual")
        if not isinstance(xdmf, XdmfWriter):
            writer.close()
    out = ctx.state_get(("uh", "eps", "RH"))
    uh = NodalFunction(out["uh"], "uh")
    if return_stats:
        stats["eps"], stats["RH"], stats["h"] = out["eps"], out["RH"], h
        return uh, stats
    return uh


def solve_kpp(domain, initial_condition=kpp_initial_condition, dt=0.01, num_steps=100, Cvel=0.5, Crv=4.0,
              bc_value=np.pi / 4, scheme="bdf2", newton_rtol=1e-4, solver="bicgstab", lin_rtol=1e-13,
              device=0, h=None, return_stats=False, xdmf=None, write_every=1, mass_rtol=0.0):
    """KPP rotating wave, BDF2-residual RV + Crank-Nicolson Newton (``KPP_exact.py``).  ``xdmf``: path of an
    XDMF time series to write (``KPP_exact.py:108-109,165``), one frame every ``write_every`` steps."""
    return _run_scalar(L.FLUX_KPP, domain, initial_condition, dt, num_steps, Cvel, Crv, "constant", bc_value,
                       scheme, newton_rtol, solver, lin_rtol, device, h, return_stats, xdmf, write_every, mass_rtol)


def solve_burgers(domain, initial_condition=burgers_initial_condition, dt=None, num_steps=None, Cvel=0.5,
                  Crv=10.0, CFL=0.5, T=0.5, scheme="bdf2", newton_rtol=1e-4, solver="bicgstab",
                  lin_rtol=1e-13, device=0, h=None, return_stats=False, xdmf=None, write_every=1, mass_rtol=0.0,
                  bc_time_lag=False):
    """2-D inviscid Burgers Riemann problem with exact Dirichlet data (``Exact_Burger_RV.py``).

    ``dt=None`` reproduces ``dt = CFL*min(h_CG)``, ``num_steps = ceil(T/dt)`` (``:105-109``).
    ``scheme="bdf1", bc_time_lag=True`` is the variant of the convergence study ``Exact_Burger_RV_conv.py``
    (BDF1 residual ``:186``, Dirichlet data one step behind ``:172-177``).
    """
    ctx = domain if isinstance(domain, Context) else Context.for_domain(domain, device=device)
    if dt is None:
        hh = ctx.nodal_h() if h is None else _interpolate(ctx, h)
        dt = CFL * float(np.min(hh))
    if num_steps is None:
        num_steps = int(np.ceil(T / dt))
    if bc_time_lag:
        # Exact_Burger_RV_conv.py:172-177 interpolates the exact solution BEFORE it advances t: step k (0-based)
        # carries the Dirichlet data of time k dt, not (k+1) dt
        bnd = ctx.boundary_dofs()
        Xb = np.zeros((3, bnd.size))
        Xb[0], Xb[1] = ctx.x[bnd, 0], ctx.x[bnd, 1]
        return _run_scalar(L.FLUX_BURGERS, ctx, initial_condition, dt, num_steps, Cvel, Crv, "user", 0.0, scheme,
                           newton_rtol, solver, lin_rtol, device, h, return_stats, xdmf, write_every, mass_rtol,
                           bc_of_step=lambda k: burgers_exact_solution(Xb, k * dt))
    return _run_scalar(L.FLUX_BURGERS, ctx, initial_condition, dt, num_steps, Cvel, Crv, "burgers_exact", 0.0,
                       scheme, newton_rtol, solver, lin_rtol, device, h, return_stats, xdmf, write_every, mass_rtol)


def solve_burgers_si(domain, initial_condition=burgers_initial_condition, dt=None, num_steps=None, Cm=0.5,
                     floor=1e-8, smooth_l=4.0, CFL=0.5, T=0.5, newton_rtol=1e-4, solver="bicgstab", lin_rtol=1e-13,
                     device=0, h=None, return_stats=False):
    """Burgers with the smoothness-indicator viscosity and the ``smooth_vector`` post-filter
    (``Code/Burgers_equation/Exact_Burger_SI.py:159-197``); ``smooth_l=0`` skips the filter.  The filter sweeps
    the nodes in the key order of ``SI.get_patch_dictionary`` like the reference."""
    ctx = domain if isinstance(domain, Context) else Context.for_domain(domain, device=device)
    h = ctx.nodal_h() if h is None else _interpolate(ctx, h)
    if dt is None:
        dt = CFL * float(np.min(h))
    if num_steps is None:
        num_steps = int(np.ceil(T / dt))
    u0 = _interpolate(ctx, initial_condition)
    ctx.state_set(uh=u0, u_n=u0, u_old=u0, u_oo=u0, RH=np.zeros(ctx.n), h=h, t=0.0)
    p = step_params(L.FLUX_BURGERS, dt, 0.0, 0.0, newton_rtol=newton_rtol, solver=solver, lin_rtol=lin_rtol,
                    bc_kind="burgers_exact")
    order = None
    if smooth_l:
        flat = np.asarray(ctx.cells).reshape(-1)
        _, first = np.unique(flat, return_index=True)
        order = flat[np.sort(first)].astype(np.int32)      # first appearance over the cells, SI.py:18-26
    stats = ctx.step_scalar_si(p, Cm, floor, smooth_l=smooth_l, smooth_order=order, n_steps=num_steps)
    out = ctx.state_get(("uh", "eps"))
    uh = NodalFunction(out["uh"], "uh")
    if return_stats:
        stats["eps"], stats["h"] = out["eps"], h
        return uh, stats
    return uh


def solve_advection(domain, initial_condition=advection_initial_condition, velocity=advection_velocity, dt=None,
                    num_steps=None, hmax=None, Cvel=0.25, Crv=1.0, CFL=0.5, T=1.0, residual_bc=False,
                    solver="bicgstab", lin_rtol=1e-13, device=0, h=None, return_stats=False, viscosity="rv"):
    """Linear advection, nodal RV, CN system rebuilt each step (``RV_node_convergence.py``).

    One GFEM step, then ``num_steps - 1`` RV steps, as the reference.  ``viscosity="none"``: plain Galerkin
    Crank-Nicolson in every step (``Code/Linear_advection/linear_advection.py:112-176``).
    """
    ctx = domain if isinstance(domain, Context) else Context.for_domain(domain, device=device)
    u0 = _interpolate(ctx, initial_condition)
    w = _interpolate(ctx, velocity)
    if dt is None:
        if hmax is None:
            raise ValueError("pass dt, or hmax (the mesh size the reference feeds gmsh, RV_node.py:82-86) to derive it")
        dt = advection_dt(w, hmax, CFL)
    if num_steps is None:
        num_steps = int(np.ceil(T / dt))
    h = ctx.nodal_h() if h is None else _interpolate(ctx, h)
    ctx.state_set(uh=u0, u_n=u0, u_old=u0, u_oo=u0, RH=np.zeros(ctx.n), h=h, w=w, t=0.0)
    p = step_params(L.FLUX_ADVECTION, dt, Cvel, Crv, scheme="bdf1", solver=solver, lin_rtol=lin_rtol,
                    bc_kind="constant", bc_value=0.0, residual_bc=residual_bc)
    if viscosity == "none":
        stats = None
        for _ in range(num_steps):     # a "first" (Galerkin) step every time; the state stays on the device
            st = ctx.step_advection(p, 1, first_gfem=True)
            if stats is None:
                stats = st
            else:
                for key in ("steps", "krylov_iterations", "kernel_launches", "spmv_launches", "assembly_launches", "device_ms"):
                    stats[key] += st[key]
                stats["time"] = st["time"]
    elif viscosity == "rv":
        stats = ctx.step_advection(p, num_steps, first_gfem=True)
    else:
        raise ValueError("viscosity must be 'rv' or 'none'")
    out = ctx.state_get(("uh", "eps", "RH"))
    uh = NodalFunction(out["uh"], "uh")
    if return_stats:
        stats["eps"], stats["RH"], stats["h"], stats["dt"] = out["eps"], out["RH"], h, dt
        return uh, stats
    return uh


# ---- compressible Euler (scheme defined by this repository, see DESIGN.md / oracle/euler.py) ----
GAMMA = 1.4


def sod_initial_condition(x, x0=1.0):
    """(rho, p) = (1, 1) left of x0, (0.125, 0.1) right of it, fluid at rest -> (N, 4) conserved state.

    gamma = 1.4 and the conserved variables follow ``Code/Compressible_euler/euler_RV.py:33,66-72``."""
    left = x[0] < x0
    rho = np.where(left, 1.0, 0.125)
    p = np.where(left, 1.0, 0.1)
    z = np.zeros_like(rho)
    return np.stack([rho, z, z, p / (GAMMA - 1.0)], axis=1)


def solve_euler(domain, initial_condition=sod_initial_condition, dt=None, num_steps=10, Cvel=0.5, Crv=4.0,
                newton_rtol=1e-4, lin_rtol=1e-13, device=0, h=None, return_stats=False):
    """4-component P1 RV Euler: BDF2 residual projection, nodal RV viscosity, Crank-Nicolson Newton.

    Every component is Dirichlet (= its initial value) on the boundary."""
    ctx = domain if isinstance(domain, Context) else Context.for_domain(domain, device=device)
    if callable(initial_condition):
        X = np.zeros((3, ctx.n))
        X[0], X[1] = ctx.x[:, 0], ctx.x[:, 1]
        U0 = np.ascontiguousarray(initial_condition(X), dtype=np.float64)
    else:
        U0 = np.ascontiguousarray(initial_condition, dtype=np.float64)
    if U0.shape != (ctx.n, 4):
        raise ValueError("the Euler state must have shape (N, 4): rho, m1, m2, E")
    h = ctx.nodal_h() if h is None else _interpolate(ctx, h)
    if dt is None:
        # no reference value exists (the reference's Euler script never ran); CFL 0.25 on the sound speed of the data
        rho, E = U0[:, 0], U0[:, 3]
        q2 = (U0[:, 1] ** 2 + U0[:, 2] ** 2) / rho ** 2
        c = np.sqrt(GAMMA * (GAMMA - 1.0) * (E / rho - 0.5 * q2))
        dt = 0.25 * float(np.min(h)) / float(np.max(np.sqrt(q2) + c))
    ctx.euler_state_set(Uh=U0, Un=U0, Uold=U0, Uoo=U0, bc_state=U0, h=h, t=0.0)
    p = step_params(L.FLUX_BURGERS, dt, Cvel, Crv, scheme="bdf2", newton_rtol=newton_rtol, lin_rtol=lin_rtol)
    stats = ctx.step_euler(p, num_steps)
    out = ctx.euler_state_get(("Uh", "eps", "R"))
    if return_stats:
        stats.update(eps=out["eps"], R=out["R"], h=h)
        return out["Uh"], stats
    return out["Uh"]


def solve_advection_rk4(domain, initial_condition=advection_initial_condition, velocity=advection_velocity, dt=None,
                        num_steps=None, hmax=None, CFL=0.5, T=1.0, lin_rtol=1e-14, device=0):
    """Galerkin linear advection with classical RK4 (``Code/Linear_advection/GFEM_RK4.py:134-218``).  Every stage
    is one residual projection on the GPU: ``k = -M_bc^{-1} int (w . grad u) v`` is ``cfem_rv_residual`` with a
    vanishing time-derivative term (``u_old = u_n``) and Dirichlet rows (``GFEM_RK4.py:127-131``)."""
    ctx = domain if isinstance(domain, Context) else Context.for_domain(domain, device=device)
    u = _interpolate(ctx, initial_condition).copy()
    w = _interpolate(ctx, velocity)
    if dt is None:
        if hmax is None:
            raise ValueError("pass dt, or hmax (the mesh size the reference feeds gmsh, RV_node.py:82-86) to derive it")
        dt = advection_dt(w, hmax, CFL)
    if num_steps is None:
        num_steps = int(np.ceil(T / dt))

    def k_of(v, guess):
        return -ctx.rv_residual("advection", "bdf1", dt, v, v, w=w, use_bc=True, R0=guess, rtol=lin_rtol)

    g = np.zeros(ctx.n)
    for _ in range(num_steps):
        k1 = k_of(u, g)
        k2 = k_of(u + 0.5 * dt * k1, -k1)
        k3 = k_of(u + 0.5 * dt * k2, -k2)
        k4 = k_of(u + dt * k3, -k3)
        u += (dt / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
        g = -k4
    return NodalFunction(u, "u_n")


# ---- (f-2) L2-error functional and convergence-rate fit -------------------------------------------------
def l2_error(domain, uh, u_ref, degree=1):
    """``sqrt(assemble_scalar((uh - u_ref)**2 * dx))``.

    ``degree=3``: ``u_ref`` (a callable of ``x`` with shape ``(3, N)``) is interpolated into P3 like the reference does
    (``Code/Burgers_equation/Exact_Burger_RV_conv.py:81-86,223``, ``Code/Linear_advection/RV_node_convergence.py:49,239``)
    -- ``Context.l2_error_p3``.  ``degree=1``: both fields in P1 (nodal values or a callable), mass-matrix norm."""
    ctx = domain if isinstance(domain, Context) else Context.for_domain(domain)
    if degree == 3:
        return ctx.l2_error_p3(_interpolate(ctx, uh), u_ref)
    if degree != 1:
        raise NotImplementedError("the error functional exists for a P1 or a P3 interpolant of the reference field")
    d = _interpolate(ctx, uh) - _interpolate(ctx, u_ref)
    Md = ctx.spmv(L.MAT_MASS, d)
    return float(np.sqrt(max(float(d @ Md), 0.0)))


def convergence_rate(hs, errors):
    """Slope of log10(error) over log10(h) (``Code/Utils/PDE_plot.py:71-73``)."""
    return float(np.polyfit(np.log10(np.asarray(hs, dtype=float)), np.log10(np.asarray(errors, dtype=float)), 1)[0])
