"""Oracle: the L2-error functional against a P3 interpolant of the exact solution.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Restates ``Code/Burgers_equation/Exact_Burger_RV_conv.py:81-86,223`` and
``Code/Linear_advection/RV_node_convergence.py:49,69-70,239``::

    W = functionspace(domain, ("Lagrange", 3)); u_exact = Function(W); u_exact.interpolate(exact_solution)
    error_L2 = sqrt(assemble_scalar((uh - u_exact)**2 * dx))

``uh`` is P1, so ``uh - u_exact`` is a P3 function on every cell and the integrand has degree 6: any rule of degree >= 6
(dolfinx picks a 12-point one, so does the reference's legacy generated kernel, ``Burger_CPP/Burger.cpp:5957-6043``)
integrates it exactly.  Here the integral is evaluated in closed form, ``e_K^T M3 e_K |K|`` with the 10 x 10 P3 mass
matrix of the reference triangle -- pinned against that generated kernel by ``tests/test_oracle_ref_kernels.py``.
"""
from __future__ import annotations

from functools import lru_cache
from math import factorial

import numpy as np

# Lagrange nodes of P3 in barycentric coordinates: 3 vertices, 2 per edge (edge opposite vertex 0, 1, 2), centroid
NODES = np.array([
    [1, 0, 0], [0, 1, 0], [0, 0, 1],
    [0, 2 / 3, 1 / 3], [0, 1 / 3, 2 / 3],
    [2 / 3, 0, 1 / 3], [1 / 3, 0, 2 / 3],
    [2 / 3, 1 / 3, 0], [1 / 3, 2 / 3, 0],
    [1 / 3, 1 / 3, 1 / 3]], dtype=np.float64)
# (i, j): node 3+k sits on edge (i, j), one third of the way from i... encoded as (near vertex, far vertex)
_EDGE = [(1, 2), (2, 1), (0, 2), (2, 0), (0, 1), (1, 0)]


def basis(lam):
    """Values of the 10 P3 Lagrange basis functions at barycentric points ``lam`` (N,3) -> (N,10)."""
    lam = np.asarray(lam, dtype=np.float64)
    out = np.empty((lam.shape[0], 10))
    for i in range(3):
        li = lam[:, i]
        out[:, i] = 0.5 * li * (3 * li - 1) * (3 * li - 2)
    for k, (i, j) in enumerate(_EDGE):
        out[:, 3 + k] = 4.5 * lam[:, i] * lam[:, j] * (3 * lam[:, i] - 1)
    out[:, 9] = 27.0 * lam[:, 0] * lam[:, 1] * lam[:, 2]
    return out


def _poly_mul(p, q):
    r = {}
    for (a, ca) in p.items():
        for (b, cb) in q.items():
            k = (a[0] + b[0], a[1] + b[1], a[2] + b[2])
            r[k] = r.get(k, 0.0) + ca * cb
    return r


def _basis_polys():
    """The basis as polynomials in (l0, l1, l2): dict exponent-triple -> coefficient."""
    lam = [{(1, 0, 0): 1.0}, {(0, 1, 0): 1.0}, {(0, 0, 1): 1.0}]
    one = {(0, 0, 0): 1.0}

    def lin(i, a, b):   # a*l_i + b
        d = {k: a * v for k, v in lam[i].items()}
        d[(0, 0, 0)] = d.get((0, 0, 0), 0.0) + b
        return d

    polys = []
    for i in range(3):
        polys.append({k: 0.5 * v for k, v in _poly_mul(_poly_mul(lam[i], lin(i, 3, -1)), lin(i, 3, -2)).items()})
    for (i, j) in _EDGE:
        polys.append({k: 4.5 * v for k, v in _poly_mul(_poly_mul(lam[i], lam[j]), lin(i, 3, -1)).items()})
    polys.append({k: 27.0 * v for k, v in _poly_mul(_poly_mul(lam[0], lam[1]), lam[2]).items()})
    del one
    return polys


@lru_cache(maxsize=1)
def mass_reference():
    """M3[a,b] = (1/|K|) int_K phi_a phi_b, from  int l0^a l1^b l2^c = 2|K| a! b! c! / (a+b+c+2)!  (exact)."""
    P = _basis_polys()
    M = np.zeros((10, 10))
    for a in range(10):
        for b in range(a, 10):
            s = 0.0
            for (e, cf) in _poly_mul(P[a], P[b]).items():
                s += cf * 2.0 * factorial(e[0]) * factorial(e[1]) * factorial(e[2]) / factorial(sum(e) + 2)
            M[a, b] = M[b, a] = s
    return M


def cell_points(x, cells):
    """Physical coordinates of the 10 P3 nodes of every cell -> (Nc,10,2)."""
    xc = np.asarray(x, dtype=np.float64)[np.asarray(cells)][:, :, :2]   # (Nc,3,2)
    return np.einsum("pk,ckd->cpd", NODES, xc)


def l2_error_p3(x, cells, uh, exact):
    """sqrt(int (uh - I3 exact)^2): ``exact`` maps points (N,2) -> values (N,), or is a ready (Nc,10) table of the
    exact solution at ``cell_points``."""
    x = np.asarray(x, dtype=np.float64)
    cells = np.asarray(cells)
    pts = cell_points(x, cells)
    ue = exact(pts.reshape(-1, 2)).reshape(-1, 10) if callable(exact) else np.asarray(exact, dtype=np.float64)
    uh3 = np.asarray(uh, dtype=np.float64)[cells] @ NODES.T          # P1 function at the P3 nodes (exact embedding)
    e = uh3 - ue
    xc = x[cells][:, :, :2]
    area = 0.5 * np.abs((xc[:, 1, 0] - xc[:, 0, 0]) * (xc[:, 2, 1] - xc[:, 0, 1])
                        - (xc[:, 1, 1] - xc[:, 0, 1]) * (xc[:, 2, 0] - xc[:, 0, 0]))
    q = np.einsum("ca,ab,cb->c", e, mass_reference(), e)
    return float(np.sqrt(np.sum(q * area)))
