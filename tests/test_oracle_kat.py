"""CPU: pin the oracle against what the reference offers.

The reference asserts nothing; its `tests/verification/*` scripts print quantities on tiny
meshes whose exact values follow analytically.  Those meshes are rebuilt here literally."""
import os

import numpy as np
import pytest

from cfem_b200 import meshes
from oracle import p1, rv, solvers as S

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_hk_test_mesh_nodal_h():
    """reference tests/verification/hk_test.py:36-38: 4 right triangles with legs 0.5 -> h_K = 0.5,
    and the L2 projection of a constant is that constant."""
    nodes = np.array([[0.0, 0.0], [0.5, 0.0], [1.0, 0.0], [0.0, 0.5], [0.5, 0.5], [1.0, 0.5]])
    conn = np.array([[0, 1, 3], [1, 3, 4], [1, 2, 4], [2, 4, 5]])
    assert np.array_equal(p1.min_edge(nodes, conn), np.full(4, 0.5))
    assert np.allclose(p1.nodal_h(nodes, conn), 0.5, rtol=0, atol=1e-15)
    # stretched variant: hypotenuse never the minimum
    nodes2 = nodes * np.array([2.0, 1.0])
    assert np.allclose(p1.min_edge(nodes2, conn), 0.5)


def test_stiffness_crossed_square():
    """reference tests/verification/stiffness.py:38: [-1,1]^2, 1x1 crossed -> centre row {4,-1,-1,-1,-1},
    corner rows {1, -1 to the centre, 0 to the adjacent corners}."""
    x, c = meshes.rectangle(1, 1, (-1, -1), (1, 1), "crossed")
    K = p1.stiffness_matrix(x, c).toarray()
    centre = int(np.argmin(np.abs(x).sum(axis=1)))
    expect = np.zeros((5, 5))
    for i in range(5):
        if i == centre:
            expect[i, :] = -1.0
            expect[i, i] = 4.0
        else:
            expect[i, i] = 1.0
            expect[i, centre] = -1.0
    assert np.allclose(K, expect, atol=1e-14)


def test_patch_test_mesh():
    """reference tests/verification/patch_test.py:15: unit square 2x2 crossed, 16 cells / 13 nodes."""
    x, c = meshes.rectangle(2, 2, diagonal="crossed")
    assert x.shape[0] == 13 and c.shape[0] == 16
    patches = p1.node_patches(c)
    centre = int(np.argmin(np.abs(x - 0.5).sum(axis=1)))
    assert len(patches[centre]) == 9           # itself + 4 mid-cell nodes + 4 edge midpoints
    corners = [i for i in range(13) if tuple(x[i]) in {(0, 0), (1, 0), (0, 1), (1, 1)}]
    assert all(len(patches[i]) == 4 for i in corners)
    mids = [i for i in range(9, 13)]
    assert all(len(patches[i]) == 5 for i in mids)
    rowptr, colidx = p1.patch_csr(c, 13)
    assert all(set(colidx[rowptr[i]:rowptr[i + 1]].tolist()) == patches[i] for i in range(13))


@pytest.mark.parametrize("mesh", ["rect", "jit", "delaunay", "ref_kpp"])
def test_p1_identities(mesh):
    if mesh == "rect":
        x, c = meshes.rectangle(11, 7, (0, 0), (2, 1))
    elif mesh == "jit":
        x, c = meshes.jittered(15, 12)
    elif mesh == "delaunay":
        x, c = meshes.delaunay(400)
    else:
        d = np.load(os.path.join(GOLD, "kpp_rv_mesh.npz"))
        x, c = d["x"], d["cells"]
    m = S.Mesh(x, c)
    n = m.n
    total = m.area.sum()
    assert abs(m.M.sum() - total) < 1e-12 * total             # 1^T M 1 = |Omega|
    lumped = p1.assemble_vector(c, np.repeat(m.area[:, None] / 3, 3, axis=1), n)
    assert np.allclose(m.M @ np.ones(n), lumped, rtol=1e-13)   # row sums = lumped mass
    K = p1.stiffness_matrix(x, c)
    assert np.abs(K @ np.ones(n)).max() < 1e-10                # K 1 = 0
    lin = 2.0 * x[:, 0] - 3.0 * x[:, 1]
    interior = np.setdiff1d(np.arange(n), m.bnd)
    assert np.abs((K @ lin)[interior]).max() < 1e-10            # linear fields are discretely harmonic
    w = S.advection_velocity(x)
    A, B = S.advection_system(m, 0.01, w)
    assert np.abs((A - m.M) @ np.ones(n)).max() < 1e-12        # C 1 = 0
    # Wathen: eigenvalues of D^-1 M in [1/2, 2] (what the Chebyshev mass solve relies on)
    if n < 600:
        d = m.M.diagonal()
        ev = np.linalg.eigvalsh((m.M.toarray() / np.sqrt(d)[:, None]) / np.sqrt(d)[None, :])
        assert ev.min() >= 0.5 - 1e-12 and ev.max() <= 2.0 + 1e-12


def test_quadrature_rules_exact():
    from math import factorial

    for deg in (2, 4, 5):
        b, w = p1.quadrature(deg)
        assert abs(w.sum() - 1) < 1e-15 and np.allclose(b.sum(axis=1), 1)
        assert len(w) == {2: 3, 4: 6, 5: 7}[deg]
        for p in range(deg + 1):
            for q in range(deg + 1 - p):
                for r in range(deg + 1 - p - q):
                    exact = 2 * factorial(p) * factorial(q) * factorial(r) / factorial(p + q + r + 2)
                    assert abs((w * b[:, 0] ** p * b[:, 1] ** q * b[:, 2] ** r).sum() - exact) < 1e-15


def test_kpp_flux_jacobian_is_derivative():
    """J (7-point rule) is the derivative of the flux vector when both use the same rule;
    with the reference's 6/7-point split they agree to quadrature accuracy."""
    x, c = meshes.jittered(6, 6, (-2, -2), (2, 2))
    m = S.Mesh(x, c)
    u = 0.3 * np.sin(x[:, 0]) + 0.2 * x[:, 1]
    for kind in ("burgers", "kpp"):
        J = p1.assemble_matrix(c, S.flux_jacobian_elements(kind, m, u), m.n)
        f = lambda v: p1.assemble_vector(c, S.flux_elements(kind, m, v), m.n)  # noqa: E731
        d = np.random.default_rng(0).normal(size=m.n)
        fd = (f(u + 1e-6 * d) - f(u - 1e-6 * d)) / 2e-6
        assert np.linalg.norm(J @ d - fd) < (1e-8 if kind == "burgers" else 1e-5) * np.linalg.norm(fd)


def test_dt_formula_matches_reference_time_stamp():
    """First <Time> of Code/Linear_advection/Data/RV/RV_node.xdmf == CFL*hmax/||w||_inf on its mesh."""
    d = np.load(os.path.join(GOLD, "rv_node_mesh.npz"))
    dt = S.advection_dt(S.advection_velocity(d["x"]), 1 / 16, CFL=0.5)
    assert dt == float(d["first_time_stamp"])
    assert repr(float(dt)).startswith(str(d["first_time_stamp_text"])[:17])


def test_epsilon_literal_equals_vectorised():
    x, c = meshes.delaunay(300)
    m = S.Mesh(x, c)
    rng = np.random.default_rng(5)
    uh, u_n, Rh = rng.normal(size=m.n), rng.normal(size=m.n), rng.normal(size=m.n)
    h = rng.uniform(0.01, 0.1, size=m.n)
    patches = p1.node_patches(c)
    for beta in (rv.beta_burgers, rv.beta_kpp):
        a = rv.epsilon_nonlinear_literal(0.5, 4.0, uh, u_n, beta, Rh, h, patches)
        b = rv.epsilon_nonlinear(0.5, 4.0, uh, u_n, beta, Rh, h, m.rowptr, m.colidx)
        assert np.array_equal(a, b)
    w = S.advection_velocity(x)
    a = rv.epsilon_linear_literal(0.25, 1.0, uh, u_n, w, Rh, h, patches)
    b = rv.epsilon_linear(0.25, 1.0, uh, u_n, w, Rh, h, m.rowptr, m.colidx)
    # np.linalg.norm (BLAS dot, possibly fused) vs sqrt(x*x+y*y): equal to 1 ulp
    assert np.allclose(a, b, rtol=4e-16, atol=0)
    # division by zero / NaN: Python min keeps the first-order branch (RV.py:83-88)
    u = np.ones(m.n)
    e = rv.epsilon_nonlinear_literal(0.5, 4.0, u, u, rv.beta_burgers, np.zeros(m.n), h, patches)
    assert np.array_equal(e, 0.5 * h * np.sqrt(2.0))


def test_literal_newton_residual_matches_direct_solve():
    x, c = meshes.rectangle(10, 10)
    m = S.Mesh(x, c)
    rng = np.random.default_rng(1)
    u = [rng.normal(size=m.n) for _ in range(3)]
    a = S.rv_residual("burgers", m, 0.01, *u, literal_newton=True, RH0=rng.normal(size=m.n))
    b = S.rv_residual("burgers", m, 0.01, *u)
    assert np.linalg.norm(a - b) < 1e-12 * np.linalg.norm(b)


def test_burgers_exact_solution_structure():
    """Exact_Burger_RV.py:37-66: far-field states and the initial condition as t -> 0+."""
    pts = np.array([[0.1, 0.9], [0.9, 0.9], [0.1, 0.1], [0.9, 0.1]])
    assert np.array_equal(S.burgers_exact(pts, 0.25), [-0.2, -1.0, 0.5, 0.8])
    g = np.random.default_rng(0).uniform(0, 1, size=(2000, 2))
    far = np.abs(g - 0.5).min(axis=1) > 0.05
    assert np.array_equal(S.burgers_exact(g, 1e-9)[far], S.burgers_initial_condition(g)[far])


def test_golden_runs_reproduce():
    """The committed golden vectors are what the oracle produces today (guards silent oracle drift)."""
    g = np.load(os.path.join(GOLD, "burgers_24x24_8steps.npz"))
    x, c = meshes.rectangle(24, 24)
    st, m, h = S.run_burgers(x, c, float(g["dt"]), 8)
    assert np.linalg.norm(st.uh - g["uh"]) <= 1e-13 * np.linalg.norm(g["uh"])
    assert list(st.newton_its) == list(g["newton_its"])


def test_advection_convergence_rate_smooth():
    """Published: L2 rate 2.25 (smooth IC).  Here: short-time check that the
    RV scheme stays second order on a smooth profile (rate > 1.7 between h=1/16 and h=1/32)."""
    errs = []
    for n in (16, 32):
        x, c = meshes.rectangle(n, n, (-1, -1), (1, 1))
        w = S.advection_velocity(x)
        dt = S.advection_dt(w, 2.0 / n) / 2
        steps = int(round(0.1 / dt))
        uh, eps, m, h = S.run_advection(x, c, dt, steps)
        th = 2 * np.pi * dt * steps
        R = np.array([[np.cos(th), np.sin(th)], [-np.sin(th), np.cos(th)]])
        exact = S.advection_initial_condition(x @ R.T)
        interior = np.linalg.norm(x, axis=1) < 0.8
        d = (uh - exact) * interior
        errs.append(np.sqrt(d @ (m.M @ d)))
    assert np.log2(errs[0] / errs[1]) > 1.7
