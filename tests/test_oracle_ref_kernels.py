"""Pins the oracle's P1 element algebra against the reference's own generated element kernels.

``oracle/build_ref.py`` cuts the five ``tabulate_tensor`` bodies out of ``/root/reference/Burger_CPP/Burger.cpp``
(FFC 2019.1 output for ``Burger.ufl``) and compiles them into ``oracle/_ref/libburger_tt.so``.  Checked here,
to 1e-14 relative, on random triangles of both orientations:

* mass  ``u v dx``                                   -> ``p1.mass_elements``
* convection  ``div(flux(u,u0)) v``  (= half of d/du of  int f'(u).grad(u) phi, f'=(u,u))
                                                     -> ``solvers.flux_jacobian_elements('burgers')``
* ``div(flux(u0,u0)) v`` = int f'(u0).grad(u0) phi   -> ``solvers.flux_elements('burgers')``
* ``eps grad u . grad v`` with a P1 coefficient      -> ``p1.stiffness_elements`` (mean of the nodal values)
* ``a_lap``                                          -> ``p1.stiffness_elements`` (unit coefficient)
* the 7-point degree-5 rule constants (Burger.cpp:5550-5563) -> ``p1.quadrature(5)``
"""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import build_ref, p1, solvers as S  # noqa: E402


@pytest.fixture(scope="module")
def forms():
    path = build_ref.build()
    if path is None:
        pytest.skip("oracle/_ref/libburger_tt.so not built and /root/reference not available")
    return build_ref.BurgerForms(path)


def _triangles(n, seed=5):
    rng = np.random.default_rng(seed)
    xy = rng.uniform(-1.0, 1.0, size=(n, 3, 2))
    area = 0.5 * ((xy[:, 1, 0] - xy[:, 0, 0]) * (xy[:, 2, 1] - xy[:, 0, 1])
                  - (xy[:, 1, 1] - xy[:, 0, 1]) * (xy[:, 2, 0] - xy[:, 0, 0]))
    keep = np.abs(area) > 0.05
    return xy[keep], rng


class _OneCell:
    """Minimal stand-in for ``solvers.Mesh`` of a single triangle (what the element routines read)."""

    def __init__(self, xy):
        self.x = np.asarray(xy, dtype=np.float64)
        self.cells = np.array([[0, 1, 2]])
        self.area, self.grad = p1.cell_geometry(self.x, self.cells)


def _hmax(xy):
    e = [np.linalg.norm(xy[a] - xy[b]) for a, b in ((0, 1), (0, 2), (1, 2))]
    return max(e)


def test_mass_convection_stiffness_matrix(forms):
    tris, rng = _triangles(60)
    assert len(tris) > 30
    worst = 0.0
    for xy in tris:
        m = _OneCell(xy)
        Me = p1.mass_elements(m.area)[0]
        k = float(rng.uniform(0.01, 0.5))
        u0 = rng.uniform(0.2, 2.0, 3)        # one sign: |u0| is then P1 and the reference's 7-point rule is exact
        alpha = rng.uniform(0.0, 1.5, 3)
        # (i) no viscosity: mass + k/2 * (1/2) d/du int f'(u).grad(u) phi
        A = forms.a(xy, u0, np.zeros(3), k)
        Jf = S.flux_jacobian_elements("burgers", m, u0)[0]
        ref = Me + 0.5 * k * 0.5 * Jf
        worst = max(worst, np.abs(A - ref).max() / np.abs(ref).max())
        # (ii) with eps = alpha * h/2 * sqrt(2) |u0|  (P1 x P1, integrated exactly): int eps = h/sqrt(2) alpha^T M u0
        A = forms.a(xy, u0, alpha, k)
        coef = _hmax(xy) / np.sqrt(2.0) * (alpha @ Me @ u0) / m.area[0]
        Ke = p1.stiffness_elements(m.area, m.grad, np.array([coef]))[0]
        ref = Me + 0.25 * k * Jf + 0.5 * k * Ke
        worst = max(worst, np.abs(A - ref).max() / np.abs(ref).max())
    assert worst < 1e-13, worst


def test_rhs_vector(forms):
    tris, rng = _triangles(60, seed=11)
    worst = 0.0
    for xy in tris:
        m = _OneCell(xy)
        Me = p1.mass_elements(m.area)[0]
        k = float(rng.uniform(0.01, 0.5))
        u0 = rng.uniform(0.2, 2.0, 3) * rng.choice([-1.0, 1.0])
        alpha = rng.uniform(0.0, 1.5, 3)
        b = forms.L(xy, u0, alpha, k)
        fe = S.flux_elements("burgers", m, u0)[0]
        coef = _hmax(xy) / np.sqrt(2.0) * (alpha @ Me @ np.abs(u0)) / m.area[0]
        Ke = p1.stiffness_elements(m.area, m.grad, np.array([coef]))[0]
        ref = Me @ u0 - 0.5 * k * fe - 0.5 * k * Ke @ u0
        worst = max(worst, np.abs(b - ref).max() / np.abs(ref).max())
    assert worst < 1e-13, worst


def test_stiffness_unit_coefficient(forms):
    tris, _ = _triangles(40, seed=3)
    for xy in tris:
        m = _OneCell(xy)
        ref = p1.stiffness_elements(m.area, m.grad)[0]
        A = forms.a_lap(xy)
        assert np.abs(A - ref).max() <= 1e-14 * np.abs(ref).max()
        # the nodal-mean rule for a P1 coefficient: K(eps) = mean(eps) K(1)
        eps = np.array([[0.3, 1.1, 0.7]])
        assert np.allclose(p1.stiffness_elements(m.area, m.grad, eps.mean(axis=1))[0], eps.mean() * A, rtol=1e-14)


def test_degree5_rule_constants(forms):
    """The 7-point rule the generated code tabulates (Burger.cpp:5550-5563) is the oracle's degree-5 rule."""
    src = open(os.path.join(build_ref.OUT_DIR, "burger_tt.cpp")).read()
    w = re.search(r"weights7\[7\] = \{([^}]*)\}", src).group(1)
    w = np.array([float(t) for t in w.split(",")])
    tab = re.search(r"FE3_C0_Q7\[1\]\[7\]\[3\] =\s*\{ \{(.*?)\} \} \};", src, re.S).group(1)
    pts = np.array([float(t) for t in re.findall(r"-?\d+\.\d+(?:[eE][-+]?\d+)?", tab)]).reshape(7, 3)
    B, W = p1.quadrature(5)
    assert abs(w.sum() - 0.5) < 1e-15 and abs(W.sum() - 1.0) < 1e-15   # FFC weights carry the reference area 1/2
    # same point set (any order) with the same weights
    key = lambda P, ww: sorted((tuple(np.round(sorted(p), 12)), round(float(v), 12)) for p, v in zip(P, ww))  # noqa: E731
    assert key(pts, 2.0 * w) == key(B, W)
    for p in pts:   # and as ordered triples, up to the basis-function numbering (1 - x - y, x, y)
        assert any(np.allclose(p, b, atol=1e-14) for b in B)


def _table(src, name, rows, cols):
    m = re.search(name + r"\[1\]\[%d\]\[%d\] =\s*\{ \{(.*?)\} \} \};" % (rows, cols), src, re.S)
    vals = [float(t) for t in re.findall(r"-?\d+\.\d+(?:[eE][-+]?\d+)?|-?\d+\.(?![\d])|-?\d+(?=[,\s}])", m.group(1))]
    return np.array(vals).reshape(rows, cols)


def test_l2_error_functional_against_p3_interpolant(forms):
    """``oracle.p3.l2_error_p3`` (closed form e^T M3 e |K|) against the reference's generated kernel for
    ``L2 = (u0 - u_ex)^2 dx`` with u_ex in P3 (Burger.ufl:40, Burger.cpp:5957-6043, 12-point degree-6 rule).

    The generated code numbers the ten P3 dofs its own way; the permutation is read off its basis-function table
    (values of the ten functions at the rule's points) instead of being assumed."""
    from oracle import p3

    src = open(os.path.join(build_ref.OUT_DIR, "burger_tt.cpp")).read()
    body = src[src.index('extern "C" void burger_tt3('):src.index('extern "C" void burger_tt4(')]
    P1tab = _table(body, "FE3_C0_Q12", 12, 3)      # P1 basis at the points == barycentric coordinates (l0, l1, l2)
    P3tab = _table(body, "FE5_C0_Q12", 12, 10)
    ours = p3.basis(P1tab)                          # our ten functions at the same points
    perm = []
    for j in range(10):                             # generated dof j is our dof perm[j]
        d = np.abs(ours - P3tab[:, [j]]).max(axis=0)
        assert d.min() < 1e-12, (j, d.min())
        perm.append(int(d.argmin()))
    assert sorted(perm) == list(range(10))
    tris, rng = _triangles(40, seed=21)
    worst = 0.0
    for xy in tris:
        u0 = rng.normal(size=3)
        ue = rng.normal(size=10)                    # in OUR dof order
        ref = forms.L2(xy, u0, ue[perm])            # int (u0 - u_ex)^2 over the cell, reference kernel
        got = p3.l2_error_p3(xy, np.array([[0, 1, 2]]), u0, ue[None, :]) ** 2
        worst = max(worst, abs(got - ref) / ref)
    assert worst < 1e-13, worst
    # a P1 function is its own P3 interpolant: zero error, and a quadratic is reproduced exactly by P3
    x = np.array([[0.0, 0.0], [1.0, 0.2], [0.3, 0.9]])
    c = np.array([[0, 1, 2]])
    lin = lambda p: 2.0 * p[:, 0] - 3.0 * p[:, 1] + 0.5  # noqa: E731
    assert p3.l2_error_p3(x, c, lin(x), lin) < 1e-15
