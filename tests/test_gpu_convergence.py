"""(f-2) The L2-error functional against a P3 interpolant and the reference's published Burgers convergence study.

* ``Context.l2_error_p3`` vs ``oracle.p3.l2_error_p3`` (itself pinned to the reference's generated L2 kernel,
  tests/test_oracle_ref_kernels.py) on an unstructured mesh, smooth and discontinuous exact fields.
* ``Code/Burgers_equation/Exact_Burger_RV_conv.py``: N = 50 / 100 / 200, CFL 0.5, T = 0.5, Cvel 0.5, CRV 10, BDF1 residual,
  Dirichlet data one step behind.  The reference publishes (figure ``Figures/RV/exact_burger_rv_conv.png``, BASELINE.md
  section 2) L2 errors of about 0.14 -> 0.105 -> 0.079 and a fitted rate of 0.43.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from cfem_b200 import Context, meshes  # noqa: E402
from cfem_b200 import solvers as GS  # noqa: E402
from oracle import p3  # noqa: E402


def test_l2_error_p3_matches_oracle():
    x, c = meshes.jittered(37, 29, (-1.0, -1.0), (2.0, 1.5))
    ctx = Context((x, c))
    rng = np.random.default_rng(3)
    uh = np.sin(2 * x[:, 0]) * np.cos(x[:, 1]) + 0.05 * rng.normal(size=ctx.n)
    smooth = lambda X: np.sin(2 * X[0]) * np.cos(X[1])                     # noqa: E731
    step = lambda X: np.where(X[0] + 0.3 * X[1] > 0.4, 1.0, -0.5)          # noqa: E731
    for f in (smooth, step):
        ref = p3.l2_error_p3(x, c, uh, lambda P: f(P.T))
        got = ctx.l2_error_p3(uh, f)
        assert abs(got - ref) <= 1e-13 * ref
        assert abs(GS.l2_error(ctx, uh, f, degree=3) - ref) <= 1e-13 * ref
    # a P1 field against itself: exactly representable in P3 -> zero error
    lin = lambda X: 2.0 * X[0] - 3.0 * X[1]                                 # noqa: E731
    assert ctx.l2_error_p3(lin(np.vstack([x.T, np.zeros(ctx.n)])), lin) < 1e-14
    ctx.close()


def _conv_case(n):
    x, c = meshes.rectangle(n, n)
    ctx = Context((x, c))
    uh, st = GS.solve_burgers(ctx, scheme="bdf1", bc_time_lag=True, CFL=0.5, T=0.5, Cvel=0.5, Crv=10.0,
                              lin_rtol=1e-11, mass_rtol=1e-11, return_stats=True)
    err = ctx.l2_error_p3(uh.x.array, lambda X: GS.burgers_exact_solution(X, 0.5))
    ctx.close()
    return err, st


def test_burgers_convergence_study_reproduces_the_published_errors():
    sizes = [50, 100, 200]
    errs, steps = [], []
    for n in sizes:
        e, st = _conv_case(n)
        errs.append(e)
        steps.append(st["steps"])
    rate = GS.convergence_rate(1.0 / np.array(sizes), errs)
    print(f"L2 errors {errs}, steps {steps}, fitted rate {rate:.3f}")
    assert all(n <= k <= n + 1 for n, k in zip(sizes, steps))   # dt = CFL min(h_CG) = 0.5 / n, steps = ceil(T / dt)
    published = [0.14, 0.105, 0.079]                      # read off the reference's figure (two significant digits)
    for e, p in zip(errs, published):
        assert abs(e - p) <= 0.06 * p, (errs, published)
    assert abs(rate - 0.43) <= 0.04


def test_convergence_case_against_the_oracle_loop():
    """The same study at N = 24 through the CPU oracle (BDF1 residual, lagged Dirichlet data): fields within 1e-10."""
    from oracle import p1, solvers as S

    n = 24
    x, c = meshes.rectangle(n, n)
    h = p1.nodal_h(x, c)
    dt = 0.5 * float(h.min())
    steps = int(np.ceil(0.5 / dt))
    bnd = p1.boundary_nodes(c, x.shape[0])
    Xb = np.vstack([x[bnd].T, np.zeros(bnd.size)])
    st, _, _ = S.run_scalar("burgers", x, c, S.burgers_initial_condition(x), dt, steps, 0.5, 10.0,
                            lambda t: GS.burgers_exact_solution(Xb, t - dt), scheme="bdf1", newton_rtol=1e-4)
    uh, gst = GS.solve_burgers((x, c), scheme="bdf1", bc_time_lag=True, CFL=0.5, T=0.5, return_stats=True)
    assert gst["steps"] == steps
    assert np.linalg.norm(uh.x.array - st.uh) <= 1e-10 * np.linalg.norm(st.uh)
    ref = p3.l2_error_p3(x, c, st.uh, lambda P: GS.burgers_exact_solution(P.T, 0.5))
    ctx = Context((x, c))
    assert abs(ctx.l2_error_p3(uh.x.array, lambda X: GS.burgers_exact_solution(X, 0.5)) - ref) <= 1e-9 * ref
    ctx.close()
