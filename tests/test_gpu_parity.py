"""GPU parity: every hot-path operator, called through the C ABI (ctypes), against
the CPU oracle on the same seeded inputs.  fp64 tolerances are written at each
assert; index work (patches, boundary dofs, CSR pattern) must match bit-exactly."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from cfem_b200 import Context, meshes, step_params, _lib as L  # noqa: E402
from cfem_b200 import solvers as GS  # noqa: E402
from oracle import p1, rv, solvers as S  # noqa: E402


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


MESHES = {
    "right_17x13": lambda: meshes.rectangle(17, 13),
    "crossed_9x9": lambda: meshes.rectangle(9, 9, (-1, -1), (1, 1), "crossed"),
    "jittered_40x33": lambda: meshes.jittered(40, 33, (-2, -2), (2, 2)),
    "delaunay_2k": lambda: meshes.delaunay(2000),
    "single_cell": lambda: (np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), np.array([[0, 1, 2]], dtype=np.int32)),
}


@pytest.fixture(scope="module", params=list(MESHES))
def case(request):
    x, c = MESHES[request.param]()
    ctx = Context((x, c))
    m = S.Mesh(x, c)
    yield request.param, x, c, ctx, m
    ctx.close()


def fields(m, seed=0):
    rng = np.random.default_rng(seed)
    x = m.x
    uh = np.sin(3 * x[:, 0]) * np.cos(2 * x[:, 1]) + 0.1 * rng.normal(size=m.n)
    u_n = uh + 0.05 * rng.normal(size=m.n)
    u_old = u_n + 0.05 * rng.normal(size=m.n)
    u_oo = u_old + 0.05 * rng.normal(size=m.n)
    return uh, u_n, u_old, u_oo, rng


def test_index_structures_bit_exact(case):
    _, x, c, ctx, m = case
    rowptr, colidx = ctx.csr_pattern()
    assert np.array_equal(rowptr, m.rowptr) and np.array_equal(colidx, m.colidx)
    assert np.array_equal(ctx.boundary_dofs(), m.bnd)
    lit = p1.node_patches(c)
    got = ctx.patch_dictionary()
    assert got == lit  # SI.get_patch_dictionary, bit-exact sets


def test_mass_matrices(case):
    _, x, c, ctx, m = case
    M = ctx.matrix(L.MAT_MASS)
    assert abs(M - m.M).max() <= 1e-15 * abs(m.M).max()
    Mbc = ctx.matrix(L.MAT_MASS_BC)
    ref = p1.apply_bc_matrix(m.M, m.bnd)
    assert abs(Mbc - ref).max() <= 1e-15 * abs(m.M).max()


def test_spmv(case):
    _, x, c, ctx, m = case
    v = np.random.default_rng(1).normal(size=m.n)
    y = ctx.spmv(L.MAT_MASS, v)
    assert rel(y, m.M @ v) < 1e-14


def test_nodal_h(case):
    _, x, c, ctx, m = case
    h = ctx.nodal_h()
    assert rel(h, p1.nodal_h(x, c)) < 1e-11


def test_stiffness(case):
    _, x, c, ctx, m = case
    eps = np.random.default_rng(2).uniform(0.1, 1.0, size=m.n)
    ctx.assemble_stiffness(eps)
    K = ctx.matrix(L.MAT_STIFFNESS)
    ref = p1.stiffness_matrix(x, c, eps)
    assert abs(K - ref).max() <= 1e-13 * abs(ref).max()


@pytest.mark.parametrize("flux", ["burgers", "kpp"])
def test_epsilon_nonlinear(case, flux):
    _, x, c, ctx, m = case
    uh, u_n, _, _, rng = fields(m)
    Rh = rng.normal(size=m.n)
    h = np.abs(rng.normal(size=m.n)) + 0.01
    beta = rv.beta_burgers if flux == "burgers" else rv.beta_kpp
    ref = rv.epsilon_nonlinear(0.5, 4.0, uh, u_n, beta, Rh, h, m.rowptr, m.colidx)
    got = ctx.rv_epsilon("nonlinear", flux, 0.5, 4.0, uh=uh, u_n=u_n, Rh=Rh, h=h)
    assert rel(got, ref) < 1e-14
    if m.n < 500:
        lit = rv.epsilon_nonlinear_literal(0.5, 4.0, uh, u_n, beta, Rh, h, p1.node_patches(c))
        assert rel(got, lit) < 1e-14


def test_epsilon_nan_inf_semantics(case):
    """n_i = 0 and Rh = 0 (0/0) or Rh > 0 (x/0): Python min keeps the first-order branch."""
    _, x, c, ctx, m = case
    u = np.full(m.n, 2.0)  # constant field: u_tilde = 0, A = 0 -> n_i = 0 everywhere
    h = np.full(m.n, 0.1)
    for Rh in (np.zeros(m.n), np.ones(m.n)):
        got = ctx.rv_epsilon("nonlinear", "burgers", 0.5, 4.0, uh=u, u_n=u, Rh=Rh, h=h)
        ref = rv.epsilon_nonlinear(0.5, 4.0, u, u, rv.beta_burgers, Rh, h, m.rowptr, m.colidx)
        assert np.all(np.isfinite(got)) and np.array_equal(got, ref)


def test_epsilon_linear_and_pointwise(case):
    _, x, c, ctx, m = case
    uh, u_n, _, _, rng = fields(m)
    Rh = rng.normal(size=m.n)
    h = np.abs(rng.normal(size=m.n)) + 0.01
    w = S.advection_velocity(x)
    ref = rv.epsilon_linear(0.25, 1.0, uh, u_n, w, Rh, h, m.rowptr, m.colidx)
    got = ctx.rv_epsilon("linear", "advection", 0.25, 1.0, uh=uh, u_n=u_n, Rh=Rh, h=h, w=w)
    assert rel(got, ref) < 1e-14
    got = ctx.rv_epsilon("pointwise", "burgers", 0.5, 4.0, uh=uh, Rh=Rh, h=h)
    assert rel(got, rv.epsilon_pointwise(0.5, 4.0, rv.beta_burgers(uh), Rh, h)) < 1e-15
    got = ctx.rv_epsilon("first_order", "kpp", 0.5, 4.0, uh=uh, h=h)
    assert rel(got, rv.epsilon_first_order(rv.beta_kpp(uh), h)) < 1e-15
    r_io = Rh.copy()
    got = ctx.rv_epsilon("linear_simple", "advection", 0.25, 1.0, u_n=u_n, Rh=r_io, h=h, w=w)
    ref, rn = rv.epsilon_linear_simple(0.25, 1.0, w, Rh, u_n, h)
    assert rel(got, ref) < 1e-14 and rel(r_io, rn) < 1e-14  # residual normalised in place


@pytest.mark.parametrize("flux,scheme", [("burgers", "bdf2"), ("kpp", "bdf2"), ("burgers", "bdf1"), ("advection", "bdf1")])
@pytest.mark.parametrize("bc", [True, False])
def test_rv_residual(case, flux, scheme, bc):
    _, x, c, ctx, m = case
    uh, u_n, u_old, u_oo, _ = fields(m)
    dt = 0.01
    w = S.advection_velocity(x) if flux == "advection" else None
    ref = S.rv_residual(flux, m, dt, u_n, u_old, u_oo, scheme=scheme, bc=bc, w=w)
    got = ctx.rv_residual(flux, scheme, dt, u_n, u_old, u_oo if scheme == "bdf2" else None, w=w, use_bc=bc)
    assert rel(got, ref) < 1e-11


@pytest.mark.parametrize("flux", ["burgers", "kpp"])
def test_cn_residual_and_jacobian(case, flux):
    _, x, c, ctx, m = case
    uh, u_n, _, _, rng = fields(m)
    eps = rng.uniform(0.0, 0.05, size=m.n)
    dt = 0.02
    g = rng.normal(size=m.bnd.size)  # inhomogeneous, != uh on the boundary: exercises lifting
    F = ctx.assemble_cn_residual(flux, dt, uh, u_n, eps, g)
    Fr = S.cn_residual(flux, m, dt, uh, u_n, eps)
    J = S.cn_jacobian(flux, m, dt, uh, eps)
    Fr = Fr + J.tocsc()[:, m.bnd] @ (g - uh[m.bnd])
    Fr[m.bnd] = uh[m.bnd] - g
    assert rel(F, Fr) < 1e-13
    ctx.assemble_cn_jacobian(flux, dt, uh, eps)
    Jg = ctx.matrix(L.MAT_SYSTEM)
    Jr = p1.apply_bc_matrix(J, m.bnd)
    assert abs(Jg - Jr).max() <= 1e-13 * abs(Jr).max()


def test_advection_system(case):
    _, x, c, ctx, m = case
    uh, u_n, _, _, rng = fields(m)
    eps = rng.uniform(0.0, 0.05, size=m.n)
    w = S.advection_velocity(x)
    dt = 0.01
    for e in (None, eps):
        b = ctx.assemble_advection(dt, w, e, u_n)
        A, B = S.advection_system(m, dt, w, e)
        br = B @ u_n
        br[m.bnd] = 0.0
        assert rel(b, br) < 1e-13
        Ag = ctx.matrix(L.MAT_SYSTEM)
        Ar = p1.apply_bc_matrix(A, m.bnd)
        assert abs(Ag - Ar).max() <= 1e-13 * abs(Ar).max()


@pytest.mark.parametrize("solver", ["pcg", "chebyshev", "bicgstab", "gmres"])
def test_krylov_vs_lu(case, solver):
    from scipy.sparse.linalg import splu

    _, x, c, ctx, m = case
    uh, u_n, _, _, rng = fields(m)
    b = rng.normal(size=m.n)
    if solver in ("pcg", "chebyshev"):
        which, A = L.MAT_MASS, m.M
    else:
        eps = rng.uniform(0.0, 0.05, size=m.n)
        ctx.assemble_cn_jacobian("burgers", 0.02, uh, eps)
        which = L.MAT_SYSTEM
        A = p1.apply_bc_matrix(S.cn_jacobian("burgers", m, 0.02, uh, eps), m.bnd)
    xg = ctx.solve(which, b, solver=solver, rtol=1e-13)
    xr = splu(A.tocsc()).solve(b)
    assert rel(xg, xr) < 1e-10
    assert ctx.last_relres <= 1e-13


def test_patch_dictionary_key_order(case):
    """keys in the reference's insertion order (first appearance over the cell loop, SI.py:18-26)."""
    _, x, c, ctx, m = case
    got = ctx.patch_dictionary()
    assert list(got.keys()) == list(p1.node_patches(c).keys()) and got.ctx is ctx


@pytest.mark.parametrize("order_kind", ["patches", "ascending", "random"])
def test_smooth_vector_matches_sequential_sweep(case, order_kind):
    """helpers.smooth_vector (helpers.py:40-50) is an in-place, order-dependent sweep; the level-scheduled
    device version must reproduce the literal loop for any sweep order.  Tolerance 1e-13: only the order of the
    <= 8 additions inside one patch sum differs (the reference leaves it to Python's set iteration)."""
    _, x, c, ctx, m = case
    uh, *_ = fields(m, seed=3)
    patches = p1.node_patches(c)
    if order_kind == "ascending":
        patches = {k: patches[k] for k in sorted(patches)}
    elif order_kind == "random":
        keys = np.random.default_rng(1).permutation(m.n)
        patches = {int(k): patches[int(k)] for k in keys}
    for l in (4.0, 6.0):
        ref = rv.smooth_vector_literal(uh.copy(), patches, l)
        got = uh.copy()
        order = None if order_kind == "ascending" else np.fromiter(patches.keys(), dtype=np.int32)
        ctx.smooth_vector(got, l, order=order)
        assert rel(got, ref) < 1e-13
        assert rel(got, uh) > 1e-3          # it did something
        # and it is NOT the Jacobi (all-old-values) version: the sweep order matters
        jac = uh.copy()
        for i, adj in patches.items():
            d = len(adj) - 1
            jac[i] = (sum(uh[j] for j in adj if j != i) + (l - 1) * d * uh[i]) / (l * d)
        if m.n > 3:
            assert rel(got, jac) > 1e-6


def test_smooth_vector_shim_and_bad_order(case):
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "conservation-fem_b200"))
    from Utils.helpers import smooth_vector
    from Utils.SI import SI

    _, x, c, ctx, m = case
    uh, *_ = fields(m, seed=4)
    patches = SI(1, ctx, 1e-8).get_patch_dictionary()
    f = GS.NodalFunction(uh.copy())
    smooth_vector(f, patches, 4)
    assert rel(f.x.array, rv.smooth_vector_literal(uh.copy(), p1.node_patches(c), 4)) < 1e-13
    with pytest.raises(L.CfemError, match="permutation|range"):
        ctx.smooth_vector(uh.copy(), 4.0, order=np.zeros(m.n, dtype=np.int32) if m.n > 1 else np.array([5], dtype=np.int32))


def test_owned_layout_state_roundtrip(case):
    """cfem_state_update_owned / cfem_state_get_owned: entry i is caller dof ordering[i] (all dofs on one GPU)."""
    _, x, c, ctx, m = case
    uh, u_n, u_old, u_oo, rng = fields(m)
    own = ctx.owned_dofs()
    assert own.size == ctx.n_owned == m.n
    ctx.state_set(uh=0 * uh, u_n=0 * uh, u_old=0 * uh, u_oo=0 * uh, RH=0 * uh, t=0.0)
    ctx.state_update_owned(uh=uh[own], u_n=u_n[own], u_old=u_old[own], u_oo=u_oo[own], t=0.25)
    g = ctx.state_get(("uh", "u_n", "u_old", "u_oo", "RH"))
    assert np.array_equal(g["uh"], uh) and np.array_equal(g["u_n"], u_n) and np.array_equal(g["u_old"], u_old)
    assert np.array_equal(g["u_oo"], u_oo) and not g["RH"].any() and g["t"] == 0.25
    o = ctx.state_get_owned(("uh", "u_oo"))
    assert np.array_equal(o["uh"], uh[own]) and np.array_equal(o["u_oo"], u_oo[own])
    with pytest.raises(ValueError):
        ctx.state_update_owned(uh=uh[:-1])


def test_si_epsilon(case):
    """(f-1) smoothness-indicator viscosity, SI.py:38-67 / 147-192."""
    _, x, c, ctx, m = case
    uh, u_n, _, _, rng = fields(m)
    h = np.abs(rng.normal(size=m.n)) + 0.01
    K = p1.stiffness_matrix(x, c)
    for use_bc in (True, False):
        bcn = m.bnd if use_bc else None
        ref, psi_ref = rv.si_epsilon(K, u_n, h, rv.beta_burgers(u_n), 1.0, 1e-8, bcn)
        got, psi = ctx.si_epsilon("burgers", 1.0, 1e-8, u_n, h, use_bc=use_bc, want_psi=True)
        assert rel(got, ref) < 1e-11 and rel(psi, psi_ref) < 1e-11
    w = S.advection_velocity(x)
    ref, _ = rv.si_epsilon(K, u_n, h, np.sqrt(w[:, 0] ** 2 + w[:, 1] ** 2), 0.5, 1e-8, m.bnd)
    got = ctx.si_epsilon("advection", 0.5, 1e-8, u_n, h, w=w)
    assert rel(got, ref) < 1e-11


def test_device_pointer_arguments(case):
    """torch is only the buffer allocator: CUDA tensors go through unchanged."""
    import torch

    _, x, c, ctx, m = case
    v = np.random.default_rng(3).normal(size=m.n)
    vd = torch.from_numpy(v).cuda()
    yd = torch.empty(m.n, dtype=torch.float64, device="cuda")
    L.check(ctx._lib.cfem_spmv(ctx._h, L.MAT_MASS, vd.data_ptr(), yd.data_ptr()))
    assert np.array_equal(yd.cpu().numpy(), ctx.spmv(L.MAT_MASS, v))


# ------------------------------------------------------------------ N-step parity
TOL_FIELD = 1e-10  # north-star tolerance: relative L2 after N steps


@pytest.mark.parametrize("solver", ["bicgstab", "gmres"])
def test_burgers_steps(solver):
    x, c = meshes.rectangle(48, 48)
    dt, n = 0.5 / 48, 12
    st, m, h = S.run_burgers(x, c, dt, n)
    uh, stats = GS.solve_burgers((x, c), dt=dt, num_steps=n, solver=solver, return_stats=True)
    assert rel(uh.x.array, st.uh) < TOL_FIELD
    assert stats["newton_iterations"] == sum(st.newton_its)
    assert rel(stats["eps"], st.eps) < 1e-8


@pytest.mark.parametrize("smooth_l", [0.0, 4.0])
def test_burgers_si_steps(smooth_l):
    """Exact_Burger_SI.py loop: SI viscosity + CN Newton (+ smooth_vector l=4, Exact_Burger_SI.py:193)."""
    x, c = meshes.jittered(36, 36) if smooth_l else meshes.rectangle(40, 40)
    dt, n = 0.5 / 40, 10
    uh_ref, eps_ref, its, m, h = S.run_burgers_si(x, c, dt, n, smooth_l=smooth_l)
    # the SI viscosity vanishes in smooth regions (nearly Galerkin, worse conditioned than the RV systems):
    # Krylov solves to 1e-14 keep the LU-based oracle within the 1e-10 bar (2e-10 at 1e-13)
    uh, stats = GS.solve_burgers_si((x, c), dt=dt, num_steps=n, smooth_l=smooth_l, lin_rtol=1e-14, return_stats=True)
    assert rel(uh.x.array, uh_ref) < TOL_FIELD
    assert stats["newton_iterations"] == sum(its)
    # alpha = |sum k_ij du| / sum |k_ij||du| cancels heavily where u is nearly flat: the 1e-10 field
    # difference of the previous step shows up as ~1e-6 in eps (in the oracle a random 1e-10 perturbation of u_n
    # moves eps by 1e-2: flat regions divide round-off by the 1e-8 floor)
    assert rel(stats["eps"], eps_ref) < 1e-4


def test_advection_gfem_steps():
    """linear_advection.py:112-176: Galerkin CN every step."""
    x, c = meshes.jittered(30, 30, (-1, -1), (1, 1))
    dt = S.advection_dt(S.advection_velocity(x), 1 / 15)
    ref, m = S.run_advection_gfem(x, c, dt, 9)
    uh, st = GS.solve_advection((x, c), dt=dt, num_steps=9, viscosity="none", return_stats=True)
    assert st["steps"] == 9 and rel(uh.x.array, ref) < TOL_FIELD
    rv_run = GS.solve_advection((x, c), dt=dt, num_steps=9)
    assert rel(rv_run.x.array, ref) > 1e-6      # the RV run is a different scheme


def test_advection_rk4_steps():
    """GFEM_RK4.py:134-218, four mass solves per step."""
    x, c = meshes.jittered(30, 30, (-1, -1), (1, 1))
    dt = S.advection_dt(S.advection_velocity(x), 1 / 15)
    ref, m = S.run_advection_rk4(x, c, dt, 8)
    uh = GS.solve_advection_rk4((x, c), dt=dt, num_steps=8)
    assert rel(uh.x.array, ref) < TOL_FIELD
    assert rel(uh.x.array, S.advection_initial_condition(x)) > 1e-2


def test_solver_writes_xdmf_series(tmp_path):
    """KPP_exact.py:108-109,165: write_mesh + write_function(uh, t) from inside the loop."""
    from cfem_b200 import io

    x, c = meshes.rectangle(24, 24)
    dt = 0.5 / 24
    ref = GS.solve_burgers((x, c), dt=dt, num_steps=7)
    path = str(tmp_path / "burgers.xdmf")
    uh, st = GS.solve_burgers((x, c), dt=dt, num_steps=7, xdmf=path, write_every=3, return_stats=True)
    assert st["steps"] == 7 and rel(uh.x.array, ref.x.array) < 1e-12
    d = io.read_xdmf(path)
    t, F = d["series"]["uh"]
    assert np.array_equal(d["x"], x) and np.array_equal(d["cells"], c)
    assert F.shape == (3, x.shape[0]) and np.allclose(t, [3 * dt, 6 * dt, 7 * dt], rtol=1e-14)
    assert np.array_equal(F[-1], uh.x.array)


def test_kpp_steps_unstructured():
    x, c = meshes.jittered(40, 40, (-2, -2), (2, 2))
    dt, n = 0.64 * 4 / 40, 10
    st, m, h = S.run_kpp(x, c, dt, n)
    uh, stats = GS.solve_kpp((x, c), dt=dt, num_steps=n, return_stats=True)
    assert rel(uh.x.array, st.uh) < TOL_FIELD
    assert stats["newton_iterations"] == sum(st.newton_its)


def test_advection_steps():
    x, c = meshes.rectangle(40, 40)
    w = S.advection_velocity(x)
    dt = S.advection_dt(w, 1 / 40)
    for rbc in (False, True):
        uh_ref, eps_ref, m, h = S.run_advection(x, c, dt, 12, residual_bc=rbc)
        uh, stats = GS.solve_advection((x, c), dt=dt, num_steps=12, residual_bc=rbc, return_stats=True)
        assert rel(uh.x.array, uh_ref) < TOL_FIELD
        assert rel(stats["eps"], eps_ref) < 1e-8


def test_euler_steps():
    """a-12: 4-component Euler RV (scheme defined here; oracle = oracle/euler.py), Sod data."""
    from oracle import euler as E

    for x, c in (meshes.rectangle(40, 20, (0, 0), (2, 1)), meshes.jittered(30, 16, (0, 0), (2, 1))):
        dt, n = 0.2 * 2 / 40, 8
        st, sol = E.run_euler(x, c, dt, n)
        U, stats = GS.solve_euler((x, c), dt=dt, num_steps=n, return_stats=True)
        assert rel(U, st["Uh"]) < TOL_FIELD
        assert stats["newton_iterations"] == sum(st["newton_its"])
        assert rel(stats["eps"], st["eps"]) < 1e-8
        assert U[:, 0].min() > 0.1 and np.all(np.isfinite(U))


def test_euler_state_round_trip_into_caller_buffers():
    """cfem_euler_state_set / _get move the (Nn,4) state through the device in the caller's numbering: what goes in
    comes out to the bit, also into caller-provided (e.g. pinned) output arrays."""
    x, c = meshes.jittered(21, 13, (0, 0), (2, 1))
    ctx = Context((x, c))
    rng = np.random.default_rng(3)
    U = rng.normal(size=(ctx.n, 4))
    ctx.euler_state_set(Uh=U, Un=U, Uold=U, Uoo=U, bc_state=U, h=ctx.nodal_h(), t=0.25)
    got = ctx.euler_state_get(("Uh",))
    assert np.array_equal(got["Uh"], U) and got["t"] == 0.25
    buf = np.full((ctx.n, 4), np.nan)
    got = ctx.euler_state_get(("Uh",), out={"Uh": buf})
    assert got["Uh"] is buf and np.array_equal(buf, U)
    with pytest.raises(ValueError):
        ctx.euler_state_get(("Uh",), out={"Uh": np.zeros((ctx.n, 3))})


def test_determinism_bitwise():
    """Atomics-free assembly and fixed-order reductions: two runs agree bit for bit."""
    x, c = meshes.jittered(32, 32, (-2, -2), (2, 2))
    a = GS.solve_kpp(Context((x, c)), dt=0.05, num_steps=4).x.array
    b = GS.solve_kpp(Context((x, c)), dt=0.05, num_steps=4).x.array
    assert np.array_equal(a, b)


def test_utils_api_surface():
    """Reference call shapes: RV(...).get_epsilon_*, SI(...).get_patch_dictionary, get_nodal_h."""
    from Utils.RV import RV
    from Utils.SI import SI
    from Utils.helpers import get_nodal_h

    x, c = meshes.rectangle(12, 12)
    domain = (x, c)
    m = S.Mesh(x, c)
    rvo = RV(0.5, 10.0, domain)
    patches = SI(1, domain, 1e-8).get_patch_dictionary()
    assert patches == p1.node_patches(c)
    h_CG = get_nodal_h(domain)
    assert rel(h_CG.x.array, p1.nodal_h(x, c)) < 1e-11
    uh, u_n, _, _, rng = fields(m)
    Rh = rng.normal(size=m.n)
    eps = rvo.get_epsilon_nonlinear(GS.NodalFunction(uh), GS.NodalFunction(u_n), lambda u: np.array([u, u]),
                                    GS.NodalFunction(Rh), h_CG, patches)
    ref = rv.epsilon_nonlinear(0.5, 10.0, uh, u_n, rv.beta_burgers, Rh, h_CG.x.array, m.rowptr, m.colidx)
    assert rel(eps.x.array, ref) < 1e-14
    eps = rvo.get_epsilon_nonlinear(uh, u_n, lambda u: np.array([np.cos(u), -np.sin(u)]), Rh, h_CG, patches)
    ref = rv.epsilon_nonlinear(0.5, 10.0, uh, u_n, rv.beta_kpp, Rh, h_CG.x.array, m.rowptr, m.colidx)
    assert rel(eps.x.array, ref) < 1e-14
    with pytest.raises(ValueError):
        rvo.get_epsilon_nonlinear(uh, u_n, lambda u: np.array([u * u, 1.0]), Rh, h_CG, patches)


def test_errors_are_loud():
    from cfem_b200 import CfemError

    x = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [5.0, 5.0]])
    with pytest.raises(CfemError):  # node 3 belongs to no cell
        Context((x, np.array([[0, 1, 2]], dtype=np.int32)))
    with pytest.raises(CfemError):  # out of range vertex
        Context((x[:3], np.array([[0, 1, 7]], dtype=np.int32)))


def test_measurement_hooks():
    """cfem_time_kernel whole-solve ids, cfem_profile_gaps: shapes and sanity (values are machine dependent)."""
    from cfem_b200 import _lib as L

    x, c = meshes.rectangle(48, 48)
    ctx = Context((x, c))
    GS.solve_burgers(ctx, dt=1e-3, num_steps=2)            # leaves an assembled system matrix behind
    ms_m, by_m = ctx.time_kernel(L.KERNEL_CHEB_ITER, "burgers", reps=2)
    ms_k, by_k = ctx.time_kernel(L.KERNEL_KRYLOV_ITER, "burgers", reps=2)
    assert ms_m > 0 and ms_k > 0 and by_m > 0 and by_k > 0
    ctx.profile_begin(20000)
    GS.solve_burgers(ctx, dt=1e-3, num_steps=2)
    gaps = ctx.profile_gaps()
    prof = ctx.profile_end()
    assert set(gaps) == set(prof) == set(ctx.PROFILE_CATEGORIES)
    assert all(v >= 0.0 for v in gaps.values()) and prof["solver"]["launches"] >= 2
