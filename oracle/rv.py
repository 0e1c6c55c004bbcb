"""Oracle: nodal residual-viscosity formulas of ``Code/Utils/RV.py``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Two restatements of each formula:

* ``*_literal``  — the reference's per-node Python loops, line for line in
  behaviour (dict patches, ``min`` with Python NaN semantics), for small cases;
* vectorised     — numpy over a CSR patch graph, for large cases; the tests
  check it against the literal form bit for bit.

``beta_norm(u)`` replaces the reference's ``velocity_field`` callable: it
returns ``||f'(u)||_2`` for an array of nodal values, computed with the same
floating-point expression the reference evaluates per node
(``np.linalg.norm(np.array(velocity_field(u)))``, ``RV.py:77-80``).
"""
from __future__ import annotations

import numpy as np


# ---- ||f'(u)||_2 per flux (same expression order as np.linalg.norm of a 2-vector)
def beta_burgers(u):
    """f'(u) = (u, u)  (``Code/Burgers_equation/Exact_Burger_RV.py:33-35``)."""
    u = np.asarray(u, dtype=np.float64)
    return np.sqrt(u * u + u * u)


def beta_kpp(u):
    """f'(u) = (cos u, -sin u)  (``Code/KPP/KPP_exact.py:55-57``)."""
    u = np.asarray(u, dtype=np.float64)
    c, s = np.cos(u), np.sin(u)
    return np.sqrt(c * c + s * s)


def _pymin(a, b):
    """Python ``min(a, b)``: returns ``a`` unless ``b < a`` (so NaN b -> a)."""
    return b if b < a else a


def absolute_term(u):
    """``np.linalg.norm(u - np.mean(u), ord=np.inf)``  (``RV.py:59,95``)."""
    u = np.asarray(u, dtype=np.float64)
    return np.max(np.abs(u - np.mean(u)))


# ------------------------------------------------------------------ literal
def epsilon_nonlinear_literal(Cvel, Crv, uh, u_n, beta_norm, Rh, h, patches):
    """``RV.get_epsilon_nonlinear`` (``Code/Utils/RV.py:56-90``)."""
    eps = np.zeros_like(np.asarray(uh, dtype=np.float64))
    A = absolute_term(uh)
    with np.errstate(divide="ignore", invalid="ignore"):
        for node, adj in patches.items():
            adj = list(adj)
            u_i = np.array([u_n[j] for j in adj])
            Rh_patch = np.array([abs(Rh[j]) for j in adj])
            beta_patch = np.array([float(beta_norm(uh[j])) for j in adj])
            u_tilde = np.max(u_i) - np.min(u_i)
            n_i = np.abs(u_tilde - A)
            Ri = np.float64(np.max(Rh_patch)) / n_i
            beta = np.max(beta_patch)
            hi = h[node]
            eps[node] = _pymin(Cvel * hi * beta, Crv * hi ** 2 * np.abs(Ri))
    return eps


def epsilon_linear_literal(Cvel, Crv, uh, u_n, w, Rh, h, patches):
    """``RV.get_epsilon_linear`` (``Code/Utils/RV.py:92-127``); w is (Nn,2)."""
    eps = np.zeros_like(np.asarray(uh, dtype=np.float64))
    A = absolute_term(uh)
    w = np.asarray(w, dtype=np.float64).reshape(-1, 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        for node, adj in patches.items():
            adj = list(adj)
            u_i = np.array([u_n[j] for j in adj])
            Rh_patch = np.array([abs(Rh[j]) for j in adj])
            beta = np.linalg.norm(w[node])  # centre node, RV.py:113-115
            u_tilde = np.max(u_i) - np.min(u_i)
            n_i = np.abs(u_tilde - A)
            Ri = np.float64(np.max(Rh_patch)) / n_i
            hi = h[node]
            eps[node] = _pymin(Cvel * hi * beta, Crv * hi ** 2 * np.abs(Ri))
    return eps


# --------------------------------------------------------------- vectorised
def _patch_reduce(ufunc, v, rowptr, colidx):
    return ufunc.reduceat(np.asarray(v)[colidx], rowptr[:-1])


def _min_pysem(a, b):
    """Elementwise Python ``min(a, b)`` semantics (NaN in b keeps a)."""
    return np.where(b < a, b, a)


def epsilon_nonlinear(Cvel, Crv, uh, u_n, beta_norm, Rh, h, rowptr, colidx):
    """Vectorised ``RV.get_epsilon_nonlinear`` on a CSR patch graph."""
    uh = np.asarray(uh, dtype=np.float64)
    A = absolute_term(uh)
    umax = _patch_reduce(np.maximum, u_n, rowptr, colidx)
    umin = _patch_reduce(np.minimum, u_n, rowptr, colidx)
    rmax = _patch_reduce(np.maximum, np.abs(Rh), rowptr, colidx)
    bmax = _patch_reduce(np.maximum, beta_norm(uh), rowptr, colidx)
    with np.errstate(divide="ignore", invalid="ignore"):
        n_i = np.abs((umax - umin) - A)
        Ri = rmax / n_i
        first = Cvel * h * bmax
        second = Crv * h ** 2 * np.abs(Ri)
        return _min_pysem(first, second)


def epsilon_linear(Cvel, Crv, uh, u_n, w, Rh, h, rowptr, colidx):
    """Vectorised ``RV.get_epsilon_linear`` (beta from the centre node)."""
    uh = np.asarray(uh, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64).reshape(-1, 2)
    A = absolute_term(uh)
    umax = _patch_reduce(np.maximum, u_n, rowptr, colidx)
    umin = _patch_reduce(np.minimum, u_n, rowptr, colidx)
    rmax = _patch_reduce(np.maximum, np.abs(Rh), rowptr, colidx)
    beta = np.sqrt(w[:, 0] * w[:, 0] + w[:, 1] * w[:, 1])
    with np.errstate(divide="ignore", invalid="ignore"):
        n_i = np.abs((umax - umin) - A)
        Ri = rmax / n_i
        return _min_pysem(Cvel * h * beta, Crv * h ** 2 * np.abs(Ri))


def epsilon_pointwise(Cvel, Crv, beta, residual, h):
    """``RV.get_epsilon`` (``RV.py:27-40``): min(Cvel h |f'|, Crv h^2 |R|)."""
    return _min_pysem(Cvel * h * beta, Crv * h ** 2 * np.abs(residual))


def epsilon_first_order(beta, h):
    """``RV.get_epsilon_1storder`` (``RV.py:42-54``)."""
    return 0.5 * h * beta


def epsilon_linear_simple(Cvel, Crv, w, residual, u_n, h):
    """``RV.get_epsilon_linear_simple`` (``RV.py:129-142``).

    Returns (eps, residual_normalised); the reference overwrites ``residual``
    in place (``RV.py:132``).
    """
    w = np.asarray(w, dtype=np.float64).reshape(-1, 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.asarray(residual, dtype=np.float64) / absolute_term(u_n)
    beta = np.sqrt(w[:, 0] * w[:, 0] + w[:, 1] * w[:, 1])
    return _min_pysem(Cvel * h * beta, Crv * h ** 2 * np.abs(r)), r


def smooth_vector_literal(u, patches, l):
    """``helpers.smooth_vector`` (``Code/Utils/helpers.py:40-50``), in place,
    order dependent (Gauss-Seidel-like sweep in dict order)."""
    for node, adj in patches.items():
        s = 0.0
        for a in adj:
            if a != node:
                s += u[a]
        d = len(adj) - 1
        u[node] = (s + (l - 1) * d * u[node]) / (l * d)
    return u


# ------------------------------------------------------------------ smoothness indicator
def sigmoid_activation(alpha, s=20.0, x0=0.5):
    """``SI.sigmoid_activation`` (``Code/Utils/SI.py:30-33``)."""
    return 1.0 / (1.0 + np.exp(-s * (alpha - x0)))


def si_epsilon(K, u_n, h, fnorm, Cm, floor, bc_nodes=None):
    """``SI.get_epsilon_nonlinear`` / ``get_epsilon_linear`` (``Code/Utils/SI.py:38-67,147-192``).

    ``K``: unit stiffness matrix (scipy CSR); with ``bc_nodes`` its Dirichlet rows/cols are identity,
    as in the reference loop (``Exact_Burger_SI.py:169-172`` assembles it with ``bcs=[bc]``).
    ``fnorm``: nodal ``||f'(u_n)||`` (nonlinear) or ``||w||`` (linear).  Returns (eps, psi)."""
    import scipy.sparse as sp

    K = sp.csr_matrix(K)
    n = K.shape[0]
    if bc_nodes is not None and len(bc_nodes):
        keep = np.ones(n)
        keep[bc_nodes] = 0.0
        D = sp.diags(keep)
        K = (D @ K @ D + sp.diags(1.0 - keep)).tocsr()
    K.sort_indices()
    rows = np.repeat(np.arange(n), np.diff(K.indptr))
    du = u_n[K.indices] - u_n[rows]
    num = np.zeros(n)
    den = np.zeros(n)
    np.add.at(num, rows, K.data * du)
    np.add.at(den, rows, np.abs(K.data) * np.abs(du))
    alpha = np.abs(num) / np.maximum(den, floor)
    psi = sigmoid_activation(alpha)
    return psi * Cm * h * fnorm, psi
