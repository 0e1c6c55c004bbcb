"""The "keeps taking a dolfinx mesh" promise of the boundary, exercised with duck-typed stand-ins for the dolfinx
objects the reference scripts pass around (no dolfinx exists in this image):

* ``Mesh``      ``.geometry.x`` (Nn, 3) float64 with z = 0, ``.geometry.dofmap`` (Nc, 3) int32
* ``Function``  ``.x.array`` (a numpy view that writes go through), ``.function_space.mesh``, ``.name``

through ``Utils.RV`` / ``Utils.SI`` / ``Utils.helpers`` exactly as ``Code/KPP/KPP_exact.py:66-137`` and
``Code/Linear_advection/RV_node.py:88-214`` call them, checked against the oracle.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from cfem_b200 import Context, meshes  # noqa: E402
from oracle import p1, rv as orv  # noqa: E402


class _Geometry:
    def __init__(self, x, cells):
        self.x = np.concatenate([np.asarray(x, dtype=np.float64), np.zeros((len(x), 1))], axis=1)   # (Nn, 3), C order
        self.dofmap = np.ascontiguousarray(cells, dtype=np.int32)


class StubMesh:
    """What ``dolfinx.mesh.Mesh`` exposes to the path."""

    def __init__(self, x, cells):
        self.geometry = _Geometry(x, cells)


class _Space:
    def __init__(self, mesh):
        self.mesh = mesh


class _Vec:
    def __init__(self, n):
        self.array = np.zeros(n)


class StubFunction:
    """``dolfinx.fem.Function`` on P1: ``.x.array``, ``.function_space.mesh``."""

    def __init__(self, mesh, values=None, name="f"):
        self.function_space = _Space(mesh)
        self.x = _Vec(mesh.geometry.x.shape[0])
        self.name = name
        if values is not None:
            self.x.array[:] = values


@pytest.fixture(scope="module")
def setup():
    x, c = meshes.jittered(23, 19, (-2.0, -2.0), (2.0, 2.0))
    return x, c, StubMesh(x, c)


def test_context_is_cached_per_mesh_object_and_not_by_id(setup):
    x, c, dom = setup
    a = Context.for_domain(dom)
    assert Context.for_domain(dom) is a                       # one context per mesh object
    other = StubMesh(*meshes.rectangle(5, 4))
    b = Context.for_domain(other)
    assert b is not a and b.n == 30 and a.n == x.shape[0]     # a different object never aliases it
    del other, b
    t1 = Context.for_domain((x, c))                           # tuples are not cached (ids of temporaries are reused)
    t2 = Context.for_domain((x, c))
    assert t1 is not t2
    t1.close(); t2.close()


def test_utils_through_stub_dolfinx_objects(setup):
    from Utils.RV import RV
    from Utils.SI import SI
    from Utils.helpers import get_nodal_h, smooth_vector

    x, c, dom = setup
    n = x.shape[0]
    rng = np.random.default_rng(4)
    h_CG = get_nodal_h(dom)                                   # helpers.py:7-38
    h_ref = p1.nodal_h(x, c)
    assert np.linalg.norm(h_CG.x.array - h_ref) < 1e-11 * np.linalg.norm(h_ref)
    si = SI(0.5, dom, 1e-8)
    patches = si.get_patch_dictionary()                       # SI.py:12-28, bit-exact incl. key order
    ref = p1.node_patches(c)
    assert list(patches.keys()) == list(ref.keys()) and all(patches[k] == ref[k] for k in ref)

    uh = StubFunction(dom, np.where(x[:, 0] ** 2 + x[:, 1] ** 2 <= 1, 3.5 * np.pi, np.pi / 4) + 0.01 * rng.normal(size=n))
    u_n = StubFunction(dom, uh.x.array + 0.02 * rng.normal(size=n))
    Rh = StubFunction(dom, rng.normal(size=n))
    kpp = lambda u: np.array([np.cos(u), -np.sin(u)])         # velocity_field callable, KPP_exact.py:55-57
    rvm = RV(0.5, 4.0, dom)
    eps = rvm.get_epsilon_nonlinear(uh, u_n, kpp, Rh, h_CG, patches)   # KPP_exact.py:139
    rowptr, colidx = p1.patch_csr(c, n)
    ref_eps = orv.epsilon_nonlinear(0.5, 4.0, uh.x.array, u_n.x.array, orv.beta_kpp, Rh.x.array, h_ref, rowptr, colidx)
    assert np.linalg.norm(eps.x.array - ref_eps) <= 1e-12 * np.linalg.norm(ref_eps)
    # patches rebuilt by hand (a plain dict, as the reference's own SI would return) are accepted when they match ...
    plain = {k: set(v) for k, v in patches.items()}
    eps2 = rvm.get_epsilon_nonlinear(uh, u_n, kpp, Rh, h_CG, plain)
    assert np.array_equal(eps2.x.array, eps.x.array)
    # ... and refused when they do not describe the mesh
    bad = {k: {k} for k in plain}
    with pytest.raises(ValueError):
        rvm.get_epsilon_nonlinear(uh, u_n, kpp, Rh, h_CG, bad)

    # linear variant with a P1 vector Function w (interleaved x.array), RV_node.py:214
    w = StubFunction(dom)
    w.x.array = np.stack([-2 * np.pi * x[:, 1], 2 * np.pi * x[:, 0]], axis=1).reshape(-1)
    eps_l = RV(0.25, 1.0, dom).get_epsilon_linear(uh, u_n, w, Rh, h_CG, patches)
    ref_l = orv.epsilon_linear(0.25, 1.0, uh.x.array, u_n.x.array, w.x.array.reshape(-1, 2), Rh.x.array, h_ref, rowptr, colidx)
    assert np.linalg.norm(eps_l.x.array - ref_l) <= 1e-12 * np.linalg.norm(ref_l)

    # in-place semantics: smooth_vector writes through the Function's array (helpers.py:40-50)
    u = StubFunction(dom, rng.normal(size=n))
    before = u.x.array.copy()
    want = before.copy()
    orv.smooth_vector_literal(want, ref, 4.0)
    smooth_vector(u, patches, 4.0)
    assert not np.array_equal(u.x.array, before)
    assert np.linalg.norm(u.x.array - want) <= 1e-13 * np.linalg.norm(want)


def test_solver_entry_points_take_the_stub_mesh(setup):
    from cfem_b200 import solvers as GS
    from oracle import solvers as S

    x, c, dom = setup
    dt = 0.64 * 4.0 / 23
    uh = GS.solve_kpp(dom, dt=dt, num_steps=3)
    st, _, _ = S.run_kpp(x, c, dt, 3)
    assert np.linalg.norm(uh.x.array - st.uh) <= 1e-10 * np.linalg.norm(st.uh)


def test_contexts_with_different_tile_footprints_coexist():
    """Kernel attributes (dynamic shared-memory opt-in, occupancy) belong to the function / device, not to one mesh:
    a context created LATER with smaller tiles must not invalidate the launches of an earlier one (regression: the
    8-GPU bench died with 'invalid argument' after its small parity contexts had lowered the opt-in)."""
    from cfem_b200 import _lib as L

    xa, ca = meshes.rectangle(70, 60)
    xb, cb = meshes.rectangle(9, 8)
    a = Context((xa, ca), order="natural")      # row-major numbering: ~140 external columns per tile
    va = np.random.default_rng(0).normal(size=a.n)
    ya = a.spmv(L.MAT_MASS, va)
    b = Context((xb, cb))                        # tiny tiles, created afterwards
    vb = np.ones(b.n)
    assert abs(b.spmv(L.MAT_MASS, vb).sum() - 1.0) < 1e-13
    assert np.array_equal(a.spmv(L.MAT_MASS, va), ya)          # the first context still launches, same bits
    h1 = a.nodal_h()
    b.close()
    assert np.array_equal(a.solve(L.MAT_MASS, ya, solver="chebyshev", rtol=1e-13), a.solve(L.MAT_MASS, ya, solver="chebyshev", rtol=1e-13))
    assert np.all(h1 > 0)
    a.close()
