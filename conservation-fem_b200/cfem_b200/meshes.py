"""Synthetic P1 triangle meshes as plain ``(x, cells)`` arrays (host, numpy).

These stand in for the meshes the reference builds with dolfinx/gmsh
(``mesh.create_rectangle`` in ``Code/Burgers_equation/Exact_Burger_RV.py:28``,
gmsh rectangle/disk in ``Code/KPP/KPP_exact.py:30-45`` and
``Code/Linear_advection/RV_node.py:32-46``); neither dolfinx nor gmsh exists in
this image.  ``x`` is (Nn, 2) float64, ``cells`` is (Nc, 3) int32 — the same two
arrays a dolfinx mesh hands over as ``geometry.x[:, :2]`` / ``geometry.dofmap``.
"""
from __future__ import annotations

import numpy as np


def rectangle(nx, ny, p0=(0.0, 0.0), p1=(1.0, 1.0), diagonal="right", rng=None):
    """Structured triangulation of a rectangle.

    ``diagonal``: 'right' (dolfinx default: cells {v0,v1,v3},{v0,v2,v3}),
    'left', 'crossed' (centre node per quad, 4 cells) or 'random' (per-quad
    coin flip between right and left, needs ``rng``).
    Node numbering is row-major ``iy*(nx+1)+ix`` (no dolfinx-style reordering).
    """
    xs = np.linspace(p0[0], p1[0], nx + 1)
    ys = np.linspace(p0[1], p1[1], ny + 1)
    X, Y = np.meshgrid(xs, ys)
    x = np.stack([X.ravel(), Y.ravel()], axis=1)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny))
    v0 = (iy * (nx + 1) + ix).ravel()
    v1 = v0 + 1
    v2 = v0 + (nx + 1)
    v3 = v2 + 1
    if diagonal == "crossed":
        mid = (nx + 1) * (ny + 1) + np.arange(nx * ny)
        xm = 0.25 * (x[v0] + x[v1] + x[v2] + x[v3])
        x = np.concatenate([x, xm], axis=0)
        cells = np.stack([
            np.stack([v0, v1, mid], 1), np.stack([v0, v2, mid], 1),
            np.stack([v1, v3, mid], 1), np.stack([v2, v3, mid], 1)], axis=1).reshape(-1, 3)
        return x, cells.astype(np.int32)
    right = np.stack([np.stack([v0, v1, v3], 1), np.stack([v0, v2, v3], 1)], axis=1)
    left = np.stack([np.stack([v0, v1, v2], 1), np.stack([v1, v2, v3], 1)], axis=1)
    if diagonal == "right":
        cells = right
    elif diagonal == "left":
        cells = left
    elif diagonal == "random":
        flip = rng.integers(0, 2, size=v0.size).astype(bool)
        cells = np.where(flip[:, None, None], left, right)
    else:
        raise ValueError(diagonal)
    return x, cells.reshape(-1, 3).astype(np.int32)


def jittered(nx, ny, p0=(0.0, 0.0), p1=(1.0, 1.0), amplitude=0.25, seed=20241118, permute=True):
    """Unstructured-like mesh (KPP variant B).

    Interior nodes of a structured grid are displaced by U(-a h, a h) per axis,
    each quad gets a random diagonal, and (``permute``) node and cell numbers
    are shuffled so nothing can rely on structured locality.
    """
    rng = np.random.default_rng(seed)
    x, cells = rectangle(nx, ny, p0, p1, diagonal="random", rng=rng)
    hx = (p1[0] - p0[0]) / nx
    hy = (p1[1] - p0[1]) / ny
    ix = np.arange((nx + 1) * (ny + 1)) % (nx + 1)
    iy = np.arange((nx + 1) * (ny + 1)) // (nx + 1)
    interior = (ix > 0) & (ix < nx) & (iy > 0) & (iy < ny)
    d = rng.uniform(-amplitude, amplitude, size=x.shape) * np.array([hx, hy])
    x = x + d * interior[:, None]
    if permute:
        x, cells = permuted(x, cells, rng)
    return x, cells


def permuted(x, cells, rng):
    """Random renumbering of nodes and cells (and of the vertex order start)."""
    n = x.shape[0]
    perm = rng.permutation(n)  # new id of old node i is inv[i]
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)
    xn = x[perm]
    cn = inv[cells]
    cn = cn[rng.permutation(cn.shape[0])]
    return xn, cn.astype(np.int32)


def delaunay(n_points, p0=(0.0, 0.0), p1=(1.0, 1.0), seed=7):
    """Genuinely unstructured Delaunay mesh of a rectangle (valence 3..10).

    Boundary points are placed regularly on the four sides; interior points are
    a jittered lattice so no sliver dominates.
    """
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    m = max(int(np.sqrt(n_points)), 3)
    t = np.linspace(0.0, 1.0, m + 1)
    bx = np.concatenate([t, t, np.zeros(m - 1), np.ones(m - 1)])
    by = np.concatenate([np.zeros(m + 1), np.ones(m + 1), t[1:-1], t[1:-1]])
    gi = (np.arange(1, m) + 0.0) / m
    GX, GY = np.meshgrid(gi, gi)
    ipts = np.stack([GX.ravel(), GY.ravel()], 1) + rng.uniform(-0.3, 0.3, size=((m - 1) ** 2, 2)) / m
    pts = np.concatenate([np.stack([bx, by], 1), ipts], axis=0)
    tri = Delaunay(pts)
    cells = tri.simplices
    # drop degenerate (collinear boundary) cells
    a = pts[cells[:, 1]] - pts[cells[:, 0]]
    b = pts[cells[:, 2]] - pts[cells[:, 0]]
    det = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
    cells = cells[np.abs(det) > 1e-14]
    x = np.asarray(p0) + pts * (np.asarray(p1) - np.asarray(p0))
    return x, cells.astype(np.int32)


def red_refine(x, cells):
    """Uniform (red) refinement: every triangle -> 4 similar triangles."""
    x = np.asarray(x, dtype=np.float64)
    c = np.asarray(cells, dtype=np.int64)
    n = x.shape[0]
    e = np.concatenate([c[:, [0, 1]], c[:, [1, 2]], c[:, [2, 0]]], axis=0)
    es = np.sort(e, axis=1)
    key = es[:, 0] * n + es[:, 1]
    uniq, inv = np.unique(key, return_inverse=True)
    mid = n + inv.reshape(3, -1).T  # (Nc,3): midpoints of edges 01,12,20
    xm = 0.5 * (x[uniq // n] + x[uniq % n])
    xn = np.concatenate([x, xm], axis=0)
    m01, m12, m20 = mid[:, 0], mid[:, 1], mid[:, 2]
    new = np.stack([
        np.stack([c[:, 0], m01, m20], 1), np.stack([m01, c[:, 1], m12], 1),
        np.stack([m20, m12, c[:, 2]], 1), np.stack([m01, m12, m20], 1)], axis=1).reshape(-1, 3)
    return xn, new.astype(np.int32)


def from_dolfinx(domain):
    """Extract (x, cells) from a dolfinx 0.9 mesh (P1 geometry).

    The reference indexes ``geometry.x`` with topology vertex ids and P1 dof
    ids interchangeably (``Code/Utils/helpers.py:20-21``); for serial P1 these
    coincide with ``geometry.dofmap`` (indexing parity).
    """
    x = np.ascontiguousarray(domain.geometry.x[:, :2], dtype=np.float64)
    cells = np.ascontiguousarray(domain.geometry.dofmap, dtype=np.int32).reshape(-1, 3)
    return x, cells


def as_mesh(domain):
    """Accept a dolfinx mesh or an ``(x, cells)`` pair."""
    if isinstance(domain, (tuple, list)) and len(domain) == 2:
        x = np.ascontiguousarray(np.asarray(domain[0], dtype=np.float64)[:, :2])
        cells = np.ascontiguousarray(np.asarray(domain[1]).reshape(-1, 3).astype(np.int32))
        return x, cells
    if hasattr(domain, "geometry"):
        return from_dolfinx(domain)
    if hasattr(domain, "x") and hasattr(domain, "cells"):
        return as_mesh((domain.x, domain.cells))
    raise TypeError("expected a dolfinx mesh or an (x, cells) pair")
