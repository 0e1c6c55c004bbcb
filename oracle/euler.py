"""Oracle: 2-D compressible Euler, 4-component P1 system with residual viscosity.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

PARITY UNPINNED — and there is nothing to pin against: the reference's
``Code/Compressible_euler/euler_RV.py`` is a non-runnable 105-line skeleton with no RV term,
no residual and no boundary condition (``LOG.md:18``: "gave up on compressible euler").  Only
these facts are taken from it: gamma = 1.4 (``:33``), a P1 vector space of conserved variables
(``:22``), conserved state (rho, m1, m2, E) (``:66-72``).  The scheme below is defined by this
repository, mirroring the scalar loop (``Code/KPP/KPP_exact.py:118-166``) component-wise:

* U = (rho, m1, m2, E),  F(U) = [m ; m (x) m / rho + p I ; (E + p) m / rho],
  p = (gamma - 1)(E - |m|^2 / (2 rho)).
* Group finite elements: F(U_h) is interpolated nodally, so
  int div F(U_h) phi_a = sum_b  Cx_ab Fx(U_b) + Cy_ab Fy(U_b),   Cd_ab = int phi_a d_d phi_b.
* Residual (BDF2):  M_bc R_k = M D_t U_k + (C . F(U_n))_k,  R = 0 on the boundary.
* Viscosity (one scalar field for all components):
      eps_i = min( Cvel h_i beta_i ,  Crv h_i^2 max_k ( max_{P(i)} |R_k| / n_{k,i} ) ),
      beta_i = max_{j in P(i)} (|u_j| + c_j)  evaluated at uh,
      n_{k,i} = | (max_{P(i)} U_k - min_{P(i)} U_k) - || U_k - mean(U_k) ||_inf |   (U_k of u_n for
      the patch range, of uh for the norm, as ``RV.get_epsilon_nonlinear``), Python-min NaN rule.
* Crank-Nicolson:  G(U) = M (U - U_n) + dt/2 [C.F(U) + C.F(U_n)] + dt/2 K_eps (U + U_n) = 0,
  Newton with dJ = M (x) I + dt/2 (Cx diag(Ax(U_b)) + Cy diag(Ay(U_b))) + dt/2 K_eps (x) I,
  dolfinx stopping rules (rtol 1e-4, atol 1e-10), all components Dirichlet (= initial data) on
  the boundary.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import splu

from . import p1, rv
from .solvers import Mesh, NewtonFailure

GAMMA = 1.4


def gradient_matrices(m: Mesh):
    """Cx, Cy with Cd[a,b] = int phi_a d_d phi_b  (= |K|/3 * grad_d phi_b per cell)."""
    out = []
    for d in (0, 1):
        Ce = np.repeat((m.area[:, None] / 3.0 * m.grad[:, :, d])[:, None, :], 3, axis=1)  # [c,a,b]
        out.append(p1.assemble_matrix(m.cells, Ce, m.n))
    return out


def pressure(U):
    rho, m1, m2, E = U[:, 0], U[:, 1], U[:, 2], U[:, 3]
    return (GAMMA - 1.0) * (E - 0.5 * (m1 * m1 + m2 * m2) / rho)


def fluxes(U):
    """Nodal fluxes Fx, Fy (N,4)."""
    rho, m1, m2, E = U[:, 0], U[:, 1], U[:, 2], U[:, 3]
    p = pressure(U)
    u, v = m1 / rho, m2 / rho
    Fx = np.stack([m1, m1 * u + p, m2 * u, (E + p) * u], axis=1)
    Fy = np.stack([m2, m1 * v, m2 * v + p, (E + p) * v], axis=1)
    return Fx, Fy


def flux_jacobians(U):
    """Ax = dFx/dU, Ay = dFy/dU at every node -> (N,4,4) each."""
    g1 = GAMMA - 1.0
    rho, m1, m2, E = U[:, 0], U[:, 1], U[:, 2], U[:, 3]
    u, v = m1 / rho, m2 / rho
    q2 = u * u + v * v
    H = (E + pressure(U)) / rho
    z, o = np.zeros_like(rho), np.ones_like(rho)
    Ax = np.array([[z, o, z, z],
                   [0.5 * g1 * q2 - u * u, (3 - GAMMA) * u, -g1 * v, g1 * o],
                   [-u * v, v, u, z],
                   [u * (0.5 * g1 * q2 - H), H - g1 * u * u, -g1 * u * v, GAMMA * u]])
    Ay = np.array([[z, z, o, z],
                   [-u * v, v, u, z],
                   [0.5 * g1 * q2 - v * v, -g1 * u, (3 - GAMMA) * v, g1 * o],
                   [v * (0.5 * g1 * q2 - H), -g1 * u * v, H - g1 * v * v, GAMMA * v]])
    return np.moveaxis(Ax, 2, 0), np.moveaxis(Ay, 2, 0)


def wave_speed(U):
    rho = U[:, 0]
    u, v = U[:, 1] / rho, U[:, 2] / rho
    c = np.sqrt(GAMMA * pressure(U) / rho)
    return np.sqrt(u * u + v * v) + c


def sod_initial_condition(x, x0=1.0):
    """(rho, p) = (1, 1) left of x0, (0.125, 0.1) right of it, fluid at rest."""
    left = x[:, 0] < x0
    rho = np.where(left, 1.0, 0.125)
    p = np.where(left, 1.0, 0.1)
    z = np.zeros_like(rho)
    return np.stack([rho, z, z, p / (GAMMA - 1.0)], axis=1)


def euler_epsilon(Cvel, Crv, Uh, Un, R, h, rowptr, colidx):
    """eps as in the module docstring.  Components whose normalised residual is NaN (0/0) are
    skipped by the max; if none is left the RV branch is +inf and the first-order value is kept."""
    red = lambda uf, v: uf.reduceat(v[colidx], rowptr[:-1])  # noqa: E731
    beta = red(np.maximum, wave_speed(Uh))
    first = Cvel * h * beta
    with np.errstate(divide="ignore", invalid="ignore"):
        Rn = np.full(Uh.shape[0], -np.inf)
        for k in range(4):
            A = rv.absolute_term(Uh[:, k])
            n_k = np.abs((red(np.maximum, Un[:, k]) - red(np.minimum, Un[:, k])) - A)
            Rk = red(np.maximum, np.abs(R[:, k])) / n_k
            Rn = np.where(Rk > Rn, Rk, Rn)
        second = Crv * h ** 2 * np.abs(Rn)
        return np.where(second < first, second, first)


class EulerRV:
    def __init__(self, x, cells):
        self.m = Mesh(x, cells)
        self.Cx, self.Cy = gradient_matrices(self.m)
        self.h = p1.nodal_h(self.m.x, self.m.cells)

    def div_flux(self, U):
        Fx, Fy = fluxes(U)
        return self.Cx @ Fx + self.Cy @ Fy

    def residual_projection(self, dt, Un, Uold, Uoo):
        m = self.m
        D = (3.0 * Un - 4.0 * Uold + Uoo) / (2.0 * dt)
        b = m.M @ D + self.div_flux(Un)
        b[m.bnd] = 0.0
        return m.mass_lu(True).solve(b)

    def step(self, st, dt, Cvel, Crv, newton_rtol=1e-4, newton_atol=1e-10, max_it=100):
        m = self.m
        n = m.n
        st["t"] += dt
        R = self.residual_projection(dt, st["Un"], st["Uold"], st["Uoo"])
        eps = euler_epsilon(Cvel, Crv, st["Uh"], st["Un"], R, self.h, m.rowptr, m.colidx)
        K = p1.stiffness_matrix(m.x, m.cells, eps)
        Un = st["Un"]
        c0 = -(m.M @ Un) + 0.5 * dt * (K @ Un) + 0.5 * dt * self.div_flux(Un)
        Sp = (m.M + 0.5 * dt * K).tocsr()
        g = st["g"]
        bnd = m.bnd

        def G(U):
            r = Sp @ U + 0.5 * dt * self.div_flux(U) + c0
            r[bnd] = U[bnd] - g[bnd]
            return r

        def Jmat(U):
            Ax, Ay = flux_jacobians(U)
            keep = np.ones(n)
            keep[bnd] = 0.0
            Dk = sp.diags(keep)
            blocks = [[None] * 4 for _ in range(4)]
            for i in range(4):
                for j in range(4):
                    B = 0.5 * dt * (self.Cx @ sp.diags(Ax[:, i, j]) + self.Cy @ sp.diags(Ay[:, i, j]))
                    if i == j:
                        B = B + Sp
                    B = Dk @ B                       # Dirichlet rows -> identity (columns stay: dU_bc = 0)
                    if i == j:
                        B = B + sp.diags(1.0 - keep)
                    blocks[i][j] = B
            return sp.bmat(blocks, format="csc")

        U = st["Uh"].copy()
        r = G(U)
        res0 = res = np.linalg.norm(r)
        converged = res < newton_atol
        it = 0
        while not converged and it < max_it:
            dU = splu(Jmat(U)).solve(r.T.reshape(-1)).reshape(4, n).T   # unknown ordering: component-major
            U = U - dU
            it += 1
            r = G(U)
            res = np.linalg.norm(r)
            converged = (res / res0 < newton_rtol) or (res < newton_atol)
        if not converged:
            raise NewtonFailure(f"Euler Newton did not converge in {it} iterations")
        st["newton_its"].append(it)
        st["Uh"], st["eps"], st["R"] = U, eps, R
        st["Uoo"], st["Uold"], st["Un"] = st["Uold"], st["Un"], U.copy()
        return st


def run_euler(x, cells, dt, num_steps, Cvel=0.5, Crv=4.0, U0=None, x0=None):
    x = np.asarray(x, dtype=np.float64)
    if U0 is None:
        U0 = sod_initial_condition(x, 0.5 * (x[:, 0].min() + x[:, 0].max()) if x0 is None else x0)
    solver = EulerRV(x, cells)
    st = {"Uh": U0.copy(), "Un": U0.copy(), "Uold": U0.copy(), "Uoo": U0.copy(), "g": U0.copy(), "t": 0.0,
          "newton_its": []}
    for _ in range(num_steps):
        solver.step(st, dt, Cvel, Crv)
    return st, solver
