"""Oracle: P1 triangle element algebra, quadrature, assembly, mesh relations.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  numpy/scipy only.

All functions take a plain mesh ``(x, cells)``: ``x`` is (Nn, 2|3) float64
node coordinates (``domain.geometry.x`` in the reference), ``cells`` is
(Nc, 3) integer connectivity (``V.dofmap.list`` == ``geometry.dofmap`` for P1).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

# --------------------------------------------------------------------------
# Quadrature on the reference triangle, barycentric points, weights sum to 1.
# FFCx picks the rule from the estimated degree of the *summed* integrand of a
# form; basix 0.9's default simplex scheme gives
# 3 / 6 / 7 points for degree 2 / 4 / 5.  The 6- and 7-point rules are the
# unique fully symmetric rules of that size and degree.
# --------------------------------------------------------------------------
_A4 = 0.4459484909159648863183292538830519883991
_B4 = 0.09157621350977074345957146340220150785433
_WA4 = 0.2233815896780114656950070084331228043703
_WB4 = 0.1099517436553218676383263249002105289631
_S15 = 15.0 ** 0.5
_A5 = (6.0 - _S15) / 21.0
_B5 = (6.0 + _S15) / 21.0
_WA5 = (155.0 - _S15) / 1200.0
_WB5 = (155.0 + _S15) / 1200.0


# Column ordering of every sparse LU in the oracle.  The matrices here have a symmetric pattern (P1 graph), for which
# minimum degree on A^T + A gives ~3x less fill and factorisation time than SuperLU's COLAMD default (1024^2 Burgers
# Jacobian: 50 s vs 150 s); it plays the role of the nested-dissection ordering PETSc's LU uses in the reference.
LU_ORDERING = "MMD_AT_PLUS_A"


def _orbit(a):
    c = 1.0 - 2.0 * a
    return [(c, a, a), (a, c, a), (a, a, c)]


def quadrature(degree: int):
    """Return (bary (nq,3), weights (nq,)) exact for polynomials of ``degree``."""
    if degree <= 2:
        pts = _orbit(1.0 / 6.0)
        w = [1.0 / 3.0] * 3
    elif degree <= 4:
        pts = _orbit(_A4) + _orbit(_B4)
        w = [_WA4] * 3 + [_WB4] * 3
    elif degree == 5:
        pts = [(1.0 / 3.0,) * 3] + _orbit(_A5) + _orbit(_B5)
        w = [9.0 / 40.0] + [_WA5] * 3 + [_WB5] * 3
    else:
        raise ValueError("oracle only needs degree <= 5")
    return np.array(pts, dtype=np.float64), np.array(w, dtype=np.float64)


# --------------------------------------------------------------------------
# Geometry
# --------------------------------------------------------------------------
def cell_geometry(x, cells):
    """Per-cell area |K| (Nc,) and P1 basis gradients (Nc, 3, 2).

    Orientation-agnostic (dolfinx uses |det J|); gradients come from the
    signed Jacobian so they are correct for clockwise cells too.
    """
    x = np.asarray(x, dtype=np.float64)[:, :2]
    c = np.asarray(cells)
    p0, p1, p2 = x[c[:, 0]], x[c[:, 1]], x[c[:, 2]]
    e1 = p1 - p0
    e2 = p2 - p0
    det = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
    g = np.empty((c.shape[0], 3, 2))
    g[:, 1, 0] = e2[:, 1] / det
    g[:, 1, 1] = -e2[:, 0] / det
    g[:, 2, 0] = -e1[:, 1] / det
    g[:, 2, 1] = e1[:, 0] / det
    g[:, 0, :] = -g[:, 1, :] - g[:, 2, :]
    return 0.5 * np.abs(det), g


def min_edge(x, cells):
    """h_K = shortest edge of each cell (``Code/Utils/helpers.py:18-24``)."""
    x = np.asarray(x, dtype=np.float64)
    c = np.asarray(cells)
    p = x[c]  # (Nc,3,dim)
    e01 = np.linalg.norm(p[:, 0] - p[:, 1], axis=1)
    e02 = np.linalg.norm(p[:, 0] - p[:, 2], axis=1)
    e12 = np.linalg.norm(p[:, 1] - p[:, 2], axis=1)
    return np.minimum(np.minimum(e01, e02), e12)


_MREF = (np.ones((3, 3)) + np.eye(3)) / 12.0  # int phi_a phi_b / |K|


# --------------------------------------------------------------------------
# Assembly helpers
# --------------------------------------------------------------------------
def assemble_matrix(cells, Ke, n):
    """Sum (Nc,3,3) element matrices into an (n,n) CSR (row = test index a)."""
    c = np.asarray(cells)
    rows = np.repeat(c, 3, axis=1).ravel()
    cols = np.tile(c, (1, 3)).ravel()
    A = sp.coo_matrix((Ke.reshape(-1), (rows, cols)), shape=(n, n)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def assemble_vector(cells, be, n):
    """Sum (Nc,3) element vectors into an (n,) vector."""
    out = np.zeros(n)
    np.add.at(out, np.asarray(cells).ravel(), be.reshape(-1))
    return out


def mass_elements(area):
    return area[:, None, None] * _MREF[None]


def stiffness_elements(area, grad, eps_cell=None):
    """K_ab = |K| * eps_mean * grad phi_a . grad phi_b  (eps P1 -> mean of 3)."""
    K = np.einsum("cad,cbd->cab", grad, grad) * area[:, None, None]
    if eps_cell is not None:
        K = K * eps_cell[:, None, None]
    return K


def convection_elements(area, grad, w_nodes_cell):
    """C_ab = int (w . grad phi_b) phi_a, w P1 vector with cell values (Nc,3,2).

    (``Code/Linear_advection/RV_node_convergence.py:110``)
    """
    # wg[c, k, b] = w_k . grad phi_b
    wg = np.einsum("ckd,cbd->ckb", w_nodes_cell, grad)
    # C_ab = sum_k wg[k,b] * M_ka
    return np.einsum("ckb,ka->cab", wg, _MREF) * area[:, None, None]


def mass_matrix(x, cells):
    area, _ = cell_geometry(x, cells)
    return assemble_matrix(cells, mass_elements(area), np.asarray(x).shape[0])


def stiffness_matrix(x, cells, eps=None):
    area, grad = cell_geometry(x, cells)
    ec = None if eps is None else np.asarray(eps)[np.asarray(cells)].mean(axis=1)
    return assemble_matrix(cells, stiffness_elements(area, grad, ec), np.asarray(x).shape[0])


# --------------------------------------------------------------------------
# Mesh relations
# --------------------------------------------------------------------------
def node_patches(cells):
    """dict node -> set(nodes sharing a cell, self included).

    Literal restatement of ``Code/Utils/SI.py:12-28`` (dict insertion order =
    first-seen order while looping cells in order).
    """
    patches: dict[int, set[int]] = {}
    for cell_nodes in np.asarray(cells):
        for node in cell_nodes:
            node = int(node)
            if node not in patches:
                patches[node] = set()
            patches[node].update(int(n) for n in cell_nodes)
    return patches


def patch_csr(cells, n):
    """Same graph as ``node_patches`` as CSR (rowptr, colidx), columns sorted."""
    c = np.asarray(cells)
    rows = np.repeat(c, 3, axis=1).ravel()
    cols = np.tile(c, (1, 3)).ravel()
    G = sp.coo_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(n, n)).tocsr()
    G.sum_duplicates()
    G.sort_indices()
    return G.indptr.astype(np.int64), G.indices.astype(np.int64)


def boundary_nodes(cells, n):
    """Sorted vertices of all boundary facets (edges owned by one cell).

    Restates ``mesh.locate_entities_boundary(domain, fdim, all-True)`` +
    ``fem.locate_dofs_topological`` (``Code/KPP/KPP_exact.py:85-89``).
    """
    c = np.asarray(cells, dtype=np.int64)
    e = np.concatenate([c[:, [0, 1]], c[:, [1, 2]], c[:, [2, 0]]], axis=0)
    e.sort(axis=1)
    key = e[:, 0] * n + e[:, 1]
    uniq, counts = np.unique(key, return_counts=True)
    b = uniq[counts == 1]
    return np.unique(np.concatenate([b // n, b % n]))


# --------------------------------------------------------------------------
# Dirichlet handling (dolfinx conventions, SURVEY.md section 8c-(2),(3))
# --------------------------------------------------------------------------
def apply_bc_matrix(A, bc_dofs):
    """Zero bc rows and columns, unit diagonal (``assemble_matrix(form, bcs)``)."""
    n = A.shape[0]
    keep = np.ones(n)
    keep[bc_dofs] = 0.0
    D = sp.diags(keep)
    Abc = (D @ A @ D).tocsr()
    Abc = Abc + sp.diags(1.0 - keep)
    return Abc.tocsr()


def nodal_h(x, cells, solve=None):
    """h_CG: L2 projection of the DG0 min-edge field onto P1.

    ``Code/Utils/helpers.py:7-38``:  M h = b,  b_i = sum_{K ni i} h_K |K| / 3.
    """
    from scipy.sparse.linalg import splu

    n = np.asarray(x).shape[0]
    area, _ = cell_geometry(x, cells)
    hk = min_edge(x, cells)
    b = assemble_vector(cells, np.repeat((hk * area / 3.0)[:, None], 3, axis=1), n)
    M = mass_matrix(x, cells)
    if solve is None:
        return splu(M.tocsc(), permc_spec=LU_ORDERING).solve(b)
    return solve(M, b)
