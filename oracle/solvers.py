"""Oracle: the reference's per-time-step loops restated on plain arrays.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  numpy/scipy, sparse LU.

Restates
* ``Code/KPP/KPP_exact.py:118-166``                      (``kpp`` flux, BDF2 residual)
* ``Code/Burgers_equation/Exact_Burger_RV.py:169-237``   (``burgers`` flux, BDF2)
* ``Code/Burgers_equation/Exact_Burger_RV_conv.py:186``  (BDF1 residual variant)
* ``Code/Linear_advection/RV_node_convergence.py:104-236`` / ``RV_node.py:206-251``
with dolfinx 0.9 ``NonlinearProblem`` / ``NewtonSolver`` semantics
(SURVEY.md section 8c, items (1)-(3)).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import splu

from . import p1, rv

# ----------------------------------------------------------------- problem data


def kpp_initial_condition(x):
    """``Code/KPP/KPP_exact.py:52-53``."""
    r2 = x[:, 0] ** 2 + x[:, 1] ** 2
    return (r2 <= 1) * 14 * np.pi / 4 + (r2 > 1) * np.pi / 4


def burgers_initial_condition(x):
    """``Code/Burgers_equation/Exact_Burger_RV.py:70-80``."""
    x0, x1 = x[:, 0], x[:, 1]
    u = np.zeros_like(x0)
    u = np.where((x0 <= 0.5) & (x1 >= 0.5), -0.2, u)
    u = np.where((x0 > 0.5) & (x1 >= 0.5), -1.0, u)
    u = np.where((x0 <= 0.5) & (x1 < 0.5), 0.5, u)
    u = np.where((x0 > 0.5) & (x1 < 0.5), 0.8, u)
    return u


def burgers_exact(x, t):
    """``Code/Burgers_equation/Exact_Burger_RV.py:37-66`` (t > 0)."""
    X, Y = x[:, 0], x[:, 1]
    u = np.zeros_like(X)
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
    l4 = X - 5 / (18 * t) * (X + t - 1 / 2) ** 2
    u = np.where(m4 & (Y > l4), -1, u)
    u = np.where(m4 & (Y <= l4), (2 * X - 1) / (2 * t), u)
    m5 = X >= (1 / 2 + 4 * t / 5)
    u = np.where(m5 & (Y > (1 / 2 - t / 10)), -1, u)
    u = np.where(m5 & (Y <= (1 / 2 - t / 10)), 0.8, u)
    return u


def advection_initial_condition(x, r0=0.25, x0_1=0.3, x0_2=0.0):
    """``Code/Linear_advection/RV_node.py:54-55``."""
    return 0.5 * (1 - np.tanh(((x[:, 0] - x0_1) ** 2 + (x[:, 1] - x0_2) ** 2) / r0 ** 2 - 1))


def advection_velocity(x):
    """``Code/Linear_advection/RV_node.py:59-60`` interpolated at the nodes -> (Nn,2)."""
    return np.stack([-2 * np.pi * x[:, 1], 2 * np.pi * x[:, 0]], axis=1)


def advection_dt(w, hmax, CFL=0.5):
    """``RV_node.py:78-85``: matrix inf-norm of the (Nn,2) velocity table."""
    w_inf_norm = np.linalg.norm(np.asarray(w).reshape(-1, 2), ord=np.inf)
    return CFL * hmax / w_inf_norm


# ---------------------------------------------------------------------- geometry


@dataclass
class Mesh:
    x: np.ndarray
    cells: np.ndarray
    n: int = 0
    area: np.ndarray = None
    grad: np.ndarray = None
    Me: np.ndarray = None
    M: sp.csr_matrix = None
    bnd: np.ndarray = None
    rowptr: np.ndarray = None
    colidx: np.ndarray = None
    _lu: dict = field(default_factory=dict)

    def __post_init__(self):
        self.x = np.asarray(self.x, dtype=np.float64)
        self.cells = np.asarray(self.cells, dtype=np.int64)
        self.n = self.x.shape[0]
        self.area, self.grad = p1.cell_geometry(self.x, self.cells)
        self.Me = p1.mass_elements(self.area)
        self.M = p1.assemble_matrix(self.cells, self.Me, self.n)
        self.bnd = p1.boundary_nodes(self.cells, self.n)
        self.rowptr, self.colidx = p1.patch_csr(self.cells, self.n)

    def mass_lu(self, bc: bool):
        key = ("M", bc)
        if key not in self._lu:
            A = p1.apply_bc_matrix(self.M, self.bnd) if bc else self.M
            self._lu[key] = splu(A.tocsc(), permc_spec=p1.LU_ORDERING)
        return self._lu[key]


# ------------------------------------------------------------------------ fluxes

_B4, _W4 = p1.quadrature(4)
_B5, _W5 = p1.quadrature(5)


def flux_elements(kind, m: Mesh, u):
    """Element vectors  int f'(u).grad(u) phi_a  -> (Nc,3)."""
    uc = u[m.cells]  # (Nc,3)
    ux = np.einsum("cb,cb->c", uc, m.grad[:, :, 0])
    uy = np.einsum("cb,cb->c", uc, m.grad[:, :, 1])
    if kind == "burgers":
        # f' = (u,u): (ux+uy) * int u phi_a      (polynomial degree 2 -> exact)
        return ((ux + uy) * m.area)[:, None] * (uc @ p1._MREF)
    if kind == "kpp":
        # f' = (cos u, -sin u); estimated degree 4 -> 6-point rule
        uq = uc @ _B4.T  # (Nc,nq)
        val = np.cos(uq) * ux[:, None] - np.sin(uq) * uy[:, None]
        return m.area[:, None] * ((val * _W4[None]) @ _B4)
    raise ValueError(kind)


def flux_jacobian_elements(kind, m: Mesh, u):
    """d/du_b of ``flux_elements`` -> (Nc,3,3) indexed [a,b]."""
    uc = u[m.cells]
    gx, gy = m.grad[:, :, 0], m.grad[:, :, 1]
    ux = np.einsum("cb,cb->c", uc, gx)
    uy = np.einsum("cb,cb->c", uc, gy)
    if kind == "burgers":
        Mu = uc @ p1._MREF  # (Nc,3)[a] = sum_c u_c M_ca
        t1 = Mu[:, :, None] * (gx + gy)[:, None, :]
        t2 = (ux + uy)[:, None, None] * p1._MREF[None]
        return (t1 + t2) * m.area[:, None, None]
    if kind == "kpp":
        # estimated degree 5 -> 7-point rule
        uq = uc @ _B5.T
        c, s = np.cos(uq), np.sin(uq)
        d = -s * ux[:, None] - c * uy[:, None]  # multiplies phi_b(q)
        # J[a,b] = sum_q w_q phi_a(q) [ d_q phi_b(q) + c_q gx_b - s_q gy_b ]
        inner = d[:, :, None] * _B5[None] + c[:, :, None] * gx[:, None, :] - s[:, :, None] * gy[:, None, :]
        J = np.einsum("q,qa,cqb->cab", _W5, _B5, inner)
        return J * m.area[:, None, None]
    raise ValueError(kind)


_BETA = {"burgers": rv.beta_burgers, "kpp": rv.beta_kpp}


# ------------------------------------------------------------------------ newton


class NewtonFailure(RuntimeError):
    pass


def newton(F_fn, J_fn, x, bc_dofs, bc_vals, rtol, atol=1e-10, max_it=50,
           criterion="residual", lu_cache=None):
    """dolfinx 0.9 ``NewtonSolver.solve`` with ``NonlinearProblem`` callbacks.

    F_fn(x) / J_fn(x): assembled residual / Jacobian WITHOUT Dirichlet handling.
    Lifting ``b -= alpha J[:,bc](g - x)`` with alpha=-1, then ``b_bc = x - g``;
    J gets bc rows/cols zeroed and unit diagonal; ``x -= dx``.
    Returns (x, iterations).  Raises ``NewtonFailure`` like dolfinx's RuntimeError.
    """
    x = x.copy()
    g = np.broadcast_to(np.asarray(bc_vals, dtype=np.float64), bc_dofs.shape)

    def F_bc(xx):
        b = F_fn(xx)
        dgx = g - xx[bc_dofs]
        if np.any(dgx != 0.0):
            J = J_fn(xx).tocsc()
            b = b + J[:, bc_dofs] @ dgx
        b[bc_dofs] = xx[bc_dofs] - g
        return b

    def solve(xx, b):
        if lu_cache is not None and "lu" in lu_cache:
            return lu_cache["lu"].solve(b)
        lu = splu(p1.apply_bc_matrix(J_fn(xx), bc_dofs).tocsc(), permc_spec=p1.LU_ORDERING)
        if lu_cache is not None:
            lu_cache["lu"] = lu
        return lu.solve(b)

    b = F_bc(x)
    it = 0
    res0 = 0.0
    converged = False
    if criterion == "residual":
        res = np.linalg.norm(b)
        converged = res < atol  # relative residual undefined at iteration 0
        res0 = res
    while not converged and it < max_it:
        dx = solve(x, b)
        x -= dx
        it += 1
        b = F_bc(x)
        if criterion == "residual":
            res = np.linalg.norm(b)
            converged = (res / res0 < rtol) or (res < atol)
        else:
            if it == 1:
                res0 = np.linalg.norm(dx)
            res = np.linalg.norm(dx)
            with np.errstate(invalid="ignore", divide="ignore"):
                converged = (res / res0 < rtol) or (res < atol)
    if not converged:
        raise NewtonFailure(f"Newton solver did not converge in {it} iterations")
    return x, it


# ----------------------------------------------------------------- scalar RV step


def rv_residual(kind, m: Mesh, dt, u_n, u_old, u_oo=None, scheme="bdf2", bc=True,
                RH0=None, literal_newton=False, w=None):
    """Projected PDE residual R_h:  M_bc R = int [D_t u + f'(u_n).grad u_n] phi.

    BDF2 (``KPP_exact.py:123-137``), BDF1 (``Exact_Burger_RV_conv.py:186``) or
    linear advection BDF1 (``RV_node.py:209-214``; kind='advection', w given).
    ``bc``: homogeneous Dirichlet on all boundary dofs (bc0), as the reference.
    ``literal_newton`` runs the reference's 'incremental' Newton on this linear
    problem (2 iterations, the 2nd a round-off no-op) from the guess ``RH0``.
    """
    if scheme == "bdf2":
        D = (3.0 * u_n - 4.0 * u_old + u_oo) / (2.0 * dt)
    elif scheme == "bdf1":
        D = (u_n - u_old) / dt
    else:
        raise ValueError(scheme)
    be = np.einsum("cab,cb->ca", m.Me, D[m.cells])
    if kind == "advection":
        Ce = p1.convection_elements(m.area, m.grad, np.asarray(w).reshape(-1, 2)[m.cells])
        be = be + np.einsum("cab,cb->ca", Ce, u_n[m.cells])
    else:
        be = be + flux_elements(kind, m, u_n)
    b = p1.assemble_vector(m.cells, be, m.n)
    lu = m.mass_lu(bc)
    if bc and not literal_newton:
        b[m.bnd] = 0.0
    if not literal_newton:
        return lu.solve(b)
    RH0 = np.zeros(m.n) if RH0 is None else RH0
    bnd = m.bnd if bc else np.zeros(0, dtype=np.int64)
    R, _ = newton(lambda r: m.M @ r - b, lambda r: m.M, RH0, bnd, 0.0,
                  rtol=1e-4, max_it=100, criterion="incremental", lu_cache={"lu": lu})
    return R


def cn_residual(kind, m: Mesh, dt, uh, u_n, eps):
    """F (``KPP_exact.py:141-145``) assembled without bc."""
    Ke = p1.stiffness_elements(m.area, m.grad, eps[m.cells].mean(axis=1))
    fe = np.einsum("cab,cb->ca", m.Me, (uh - u_n)[m.cells])
    fe += 0.5 * dt * (flux_elements(kind, m, uh) + flux_elements(kind, m, u_n))
    fe += 0.5 * dt * np.einsum("cab,cb->ca", Ke, (uh + u_n)[m.cells])
    return p1.assemble_vector(m.cells, fe, m.n)


def cn_jacobian(kind, m: Mesh, dt, uh, eps):
    """J = dF/duh (eps is a coefficient, not differentiated)."""
    Ke = p1.stiffness_elements(m.area, m.grad, eps[m.cells].mean(axis=1))
    Je = m.Me + 0.5 * dt * flux_jacobian_elements(kind, m, uh) + 0.5 * dt * Ke
    return p1.assemble_matrix(m.cells, Je, m.n)


@dataclass
class ScalarState:
    uh: np.ndarray
    u_n: np.ndarray
    u_old: np.ndarray
    u_oo: np.ndarray
    RH: np.ndarray
    t: float = 0.0
    newton_its: list = field(default_factory=list)
    eps: np.ndarray = None


def scalar_rv_step(kind, m: Mesh, st: ScalarState, dt, Cvel, Crv, h, bc_vals,
                   scheme="bdf2", newton_rtol=1e-4, newton_max_it=100, literal_newton=False):
    """One pass of the loop body ``KPP_exact.py:119-161`` / ``Exact_Burger_RV.py:170-224``."""
    st.t += dt
    g = bc_vals(st.t) if callable(bc_vals) else bc_vals
    st.RH = rv_residual(kind, m, dt, st.u_n, st.u_old, st.u_oo, scheme=scheme, bc=True,
                        RH0=st.RH, literal_newton=literal_newton)
    st.eps = rv.epsilon_nonlinear(Cvel, Crv, st.uh, st.u_n, _BETA[kind], st.RH, h, m.rowptr, m.colidx)
    eps, u_n = st.eps, st.u_n
    st.uh, its = newton(lambda u: cn_residual(kind, m, dt, u, u_n, eps),
                        lambda u: cn_jacobian(kind, m, dt, u, eps),
                        st.uh, m.bnd, g, rtol=newton_rtol, max_it=newton_max_it)
    st.newton_its.append(its)
    st.u_oo = st.u_old.copy()
    st.u_old = st.u_n.copy()
    st.u_n = st.uh.copy()
    return st


def run_scalar(kind, x, cells, u0, dt, num_steps, Cvel, Crv, bc_vals, h=None, **kw):
    """Whole scalar run (KPP / Burgers) from the initial condition ``u0``."""
    m = Mesh(x, cells)
    if h is None:
        h = p1.nodal_h(m.x, m.cells)
    st = ScalarState(u0.copy(), u0.copy(), u0.copy(), u0.copy(), np.zeros(m.n))
    for _ in range(num_steps):
        scalar_rv_step(kind, m, st, dt, Cvel, Crv, h, bc_vals, **kw)
    return st, m, h


def run_kpp(x, cells, dt, num_steps, Cvel=0.5, Crv=4.0, **kw):
    """``Code/KPP/KPP_exact.py`` (Cvel/CRV ``:77-78``, bc = pi/4 ``:88``)."""
    x = np.asarray(x, dtype=np.float64)
    return run_scalar("kpp", x, cells, kpp_initial_condition(x), dt, num_steps, Cvel, Crv,
                      np.pi / 4, newton_rtol=1e-4, **kw)


def run_burgers(x, cells, dt, num_steps, Cvel=0.5, Crv=10.0, **kw):
    """``Code/Burgers_equation/Exact_Burger_RV.py`` (exact Dirichlet data each step)."""
    x = np.asarray(x, dtype=np.float64)
    bnd = p1.boundary_nodes(cells, x.shape[0])
    return run_scalar("burgers", x, cells, burgers_initial_condition(x), dt, num_steps, Cvel, Crv,
                      lambda t: burgers_exact(x[bnd], t), newton_rtol=1e-4, **kw)


def run_burgers_si(x, cells, dt, num_steps, Cm=0.5, floor=1e-8, smooth_l=4.0):
    """``Code/Burgers_equation/Exact_Burger_SI.py:159-197``: per step the SI viscosity from the bc'd unit
    stiffness matrix and ``u_n`` (``SI.py:38-67``), the Crank-Nicolson Newton solve with exact Dirichlet data,
    then ``smooth_vector(uh, node_patches, l)`` (literal in-place sweep in the patches' key order;
    ``smooth_l = 0`` skips it)."""
    x = np.asarray(x, dtype=np.float64)
    m = Mesh(x, cells)
    h = p1.nodal_h(m.x, m.cells)
    K = p1.stiffness_matrix(m.x, m.cells)
    patches = p1.node_patches(m.cells)
    uh = burgers_initial_condition(x)
    u_n = uh.copy()
    t, its_all = 0.0, []
    for _ in range(num_steps):
        t += dt
        g = burgers_exact(x[m.bnd], t)
        eps, _ = rv.si_epsilon(K, u_n, h, rv.beta_burgers(u_n), Cm, floor, bc_nodes=m.bnd)
        un = u_n
        uh, its = newton(lambda u: cn_residual("burgers", m, dt, u, un, eps),
                         lambda u: cn_jacobian("burgers", m, dt, u, eps),
                         uh, m.bnd, g, rtol=1e-4, max_it=100)
        its_all.append(its)
        if smooth_l:
            rv.smooth_vector_literal(uh, patches, smooth_l)
        u_n = uh.copy()
    return uh, eps, its_all, m, h


# --------------------------------------------------------------- linear advection


def advection_system(m: Mesh, dt, w, eps=None):
    """(A, B):  A = M + dt/2 C + dt/2 K_eps ,  B = M - dt/2 C - dt/2 K_eps  (no bc).

    ``RV_node_convergence.py:110-111`` (eps=None) and ``:207-208``.
    """
    Ce = p1.convection_elements(m.area, m.grad, np.asarray(w).reshape(-1, 2)[m.cells])
    Se = 0.5 * dt * Ce
    if eps is not None:
        Se = Se + 0.5 * dt * p1.stiffness_elements(m.area, m.grad, eps[m.cells].mean(axis=1))
    return (p1.assemble_matrix(m.cells, m.Me + Se, m.n),
            p1.assemble_matrix(m.cells, m.Me - Se, m.n))


def advection_solve(m: Mesh, A, B, u_n, g=0.0):
    """assemble_vector + apply_lifting + set_bc + LU solve (``:224-233``)."""
    b = B @ u_n
    gv = np.full(m.bnd.shape, g, dtype=np.float64)
    if np.any(gv != 0.0):
        b = b - A.tocsc()[:, m.bnd] @ gv
    b[m.bnd] = gv
    return splu(p1.apply_bc_matrix(A, m.bnd).tocsc(), permc_spec=p1.LU_ORDERING).solve(b)


def run_advection(x, cells, dt, num_steps, Cvel=0.25, Crv=1.0, u0=None, w=None, h=None,
                  residual_bc=False):
    """``Code/Linear_advection/RV_node_convergence.py:104-236``.

    One GFEM Crank-Nicolson step, then ``num_steps - 1`` RV steps.
    ``residual_bc=True`` gives ``RV_node.py:213`` (residual projected with bc).
    """
    x = np.asarray(x, dtype=np.float64)
    m = Mesh(x, cells)
    u0 = advection_initial_condition(x) if u0 is None else u0
    w = advection_velocity(x) if w is None else w
    h = p1.nodal_h(m.x, m.cells) if h is None else h
    u_n, u_old = u0.copy(), u0.copy()
    A, B = advection_system(m, dt, w)
    uh = advection_solve(m, A, B, u_n)
    u_n = uh.copy()
    eps = np.zeros(m.n)
    for _ in range(num_steps - 1):
        Rh = rv_residual("advection", m, dt, u_n, u_old, scheme="bdf1", bc=residual_bc, w=w)
        eps = rv.epsilon_linear(Cvel, Crv, uh, u_n, w, Rh, h, m.rowptr, m.colidx)
        A, B = advection_system(m, dt, w, eps)
        uh = advection_solve(m, A, B, u_n)
        u_old = u_n.copy()
        u_n = uh.copy()
    return uh, eps, m, h


def run_advection_gfem(x, cells, dt, num_steps, u0=None, w=None):
    """``Code/Linear_advection/linear_advection.py:112-176``: Galerkin Crank-Nicolson, constant system matrix."""
    x = np.asarray(x, dtype=np.float64)
    m = Mesh(x, cells)
    u = advection_initial_condition(x) if u0 is None else np.array(u0, dtype=np.float64)
    w = advection_velocity(x) if w is None else w
    A, B = advection_system(m, dt, w)
    for _ in range(num_steps):
        u = advection_solve(m, A, B, u)
    return u, m


def run_advection_rk4(x, cells, dt, num_steps, u0=None, w=None):
    """``Code/Linear_advection/GFEM_RK4.py:134-218``: Galerkin linear advection, classical RK4 in time; every
    stage solves ``M_bc k = -int (w . grad u) v`` with homogeneous Dirichlet rows (LU in the reference)."""
    x = np.asarray(x, dtype=np.float64)
    m = Mesh(x, cells)
    u = advection_initial_condition(x) if u0 is None else np.array(u0, dtype=np.float64)
    w = advection_velocity(x) if w is None else w
    C = p1.assemble_matrix(m.cells, p1.convection_elements(m.area, m.grad, np.asarray(w).reshape(-1, 2)[m.cells]), m.n)
    lu = splu(p1.apply_bc_matrix(m.M, m.bnd).tocsc(), permc_spec=p1.LU_ORDERING)

    def k_of(v):
        b = -(C @ v)
        b[m.bnd] = 0.0
        return lu.solve(b)

    for _ in range(num_steps):
        k1 = k_of(u)
        k2 = k_of(u + 0.5 * dt * k1)
        k3 = k_of(u + 0.5 * dt * k2)
        k4 = k_of(u + dt * k3)
        u = u + (dt / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
    return u, m


# ------------------------------------------------------------------ functionals


def l2_error_nodal(m: Mesh, uh, u_ref):
    """sqrt(int (uh - I_h u_ref)^2) with both fields P1 (mass-matrix norm)."""
    d = uh - u_ref
    return float(np.sqrt(d @ (m.M @ d)))


# ------------------------------------------------------------------ the reference's stored dolfinx series
# Restatements of the three scripts whose XDMF/HDF5 output the reference keeps under
# ``Code/Linear_advection/Data`` (unit disk, gmsh h = 1/16, 1,011 nodes, 285 frames each).  They differ
# from ``run_advection`` only in how the nodal viscosity is formed; the tests compare every stored frame
# with these (``tests/test_oracle_pinned.py``) -- this is what pins the oracle against dolfinx itself.


def discontinuous_cylinder(x, r0=0.25, x0_1=0.3, x0_2=0.0):
    """``tests/eps_func.py:44-45`` (also ``RV_cell.py:44-45``, ``smoothness.py`` "Discont. IC")."""
    x = np.asarray(x, dtype=np.float64)
    return ((x[:, 0] - x0_1) ** 2 + (x[:, 1] - x0_2) ** 2 <= r0 ** 2).astype(np.float64)


def si_alpha(Kw, u, floor=1e-8):
    """alpha_i = |sum_j k_ij (u_j - u_i)| / max(sum_j |k_ij| |u_j - u_i|, floor) with the weights k_ij
    read from the CSR matrix ``Kw`` (``smoothness_old_convergence.py:228-246``, ``SI.py:160-185``)."""
    Kw = Kw.tocsr()
    Kw.sort_indices()
    n = Kw.shape[0]
    rows = np.repeat(np.arange(n), np.diff(Kw.indptr))
    du = u[Kw.indices] - u[rows]
    num = np.zeros(n)
    den = np.zeros(n)
    np.add.at(num, rows, Kw.data * du)
    np.add.at(den, rows, np.abs(Kw.data) * np.abs(du))
    return np.abs(num) / np.maximum(den, floor)


def run_advection_stored(x, cells, variant, num_steps=None, hmax=1 / 16, Cvel=0.25, Crv=1.0, Cm=0.05):
    """All frames ``uh(t_k)``, k = 1..num_steps, of one of the stored runs.

    variant
      ``"eps_func"``  ``tests/eps_func.py:166-249`` -> ``Data/RV/RV_node.h5``: one GFEM step, then BDF1
                      residual projected *without* bc, divided by ``max(u_n - mean u_n)``, pointwise
                      ``min(Cvel h |w|, Crv h^2 |R|)``.
      ``"rv_cell"``   ``Code/Linear_advection/RV_cell.py:166-231`` -> ``Data/RV/RV_cell.h5``: same residual
                      (the stored run predates the ``bcs=[bc]`` in line 173); per cell
                      ``min(Cvel h_K max|w|, Crv h_K^2 max|R|)`` written to the cell's three dofs in cell
                      order, so each node keeps the value of its highest-numbered cell.
      ``"si_old"``    ``smoothness_old_convergence.py:184-253`` -> ``Data/SI/smoothness.h5``:
                      ``eps = alpha Cm h |w|`` with no activation, Cm = 0.05; the script rebinds the name
                      ``A`` from the bc'd unit stiffness matrix to the step's system matrix (line 181 vs
                      the loop body), so from the second SI step on the weights ``A.getValue(i, j)`` are
                      the previous step's Crank-Nicolson matrix.  Restated as run.
    """
    x = np.asarray(x, dtype=np.float64)
    m = Mesh(x, cells)
    w = advection_velocity(x)
    wn = np.sqrt(w[:, 0] * w[:, 0] + w[:, 1] * w[:, 1])
    dt = advection_dt(w, hmax)
    num_steps = int(np.ceil(1.0 / dt)) if num_steps is None else num_steps
    h = p1.nodal_h(m.x, m.cells)
    hk = p1.min_edge(m.x, m.cells)
    u_n = discontinuous_cylinder(x)
    u_old = u_n.copy()
    A, B = advection_system(m, dt, w)
    uh = advection_solve(m, A, B, u_n)
    u_n = uh.copy()
    frames = [uh.copy()]
    if variant == "si_old":
        Kw = p1.apply_bc_matrix(p1.stiffness_matrix(m.x, m.cells), m.bnd)
    for _ in range(num_steps - 1):
        if variant == "si_old":
            eps = si_alpha(Kw, u_n) * Cm * h * wn
        else:
            Rh = rv_residual("advection", m, dt, u_n, u_old, scheme="bdf1", bc=False, w=w)
            Rh = Rh / np.max(u_n - np.mean(u_n))
            if variant == "eps_func":
                eps = rv.epsilon_pointwise(Cvel, Crv, wn, Rh, h)
            elif variant == "rv_cell":
                ek = np.minimum(Cvel * hk * wn[m.cells].max(axis=1), Crv * hk ** 2 * np.abs(Rh)[m.cells].max(axis=1))
                eps = np.zeros(m.n)
                for k in range(len(m.cells)):       # last writer wins (RV_cell.py:190-192)
                    eps[m.cells[k]] = ek[k]
            else:
                raise ValueError(variant)
        A, B = advection_system(m, dt, w, eps)
        if variant == "si_old":
            Kw = p1.apply_bc_matrix(A, m.bnd)
        uh = advection_solve(m, A, B, u_n)
        u_old = u_n.copy()
        u_n = uh.copy()
        frames.append(uh.copy())
    return np.array(frames), m, dt
