"""GPU: (1) the committed golden vectors (tests/golden, made by make_golden.py from the oracle on
the reference's own meshes); (2) BASELINE-size runs checked through size-independent properties."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from cfem_b200 import Context, meshes, step_params, _lib as L  # noqa: E402
from cfem_b200 import solvers as GS  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-10  # north-star tolerance, relative L2 of the field after N steps


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


def test_golden_burgers():
    g = np.load(os.path.join(GOLD, "burgers_24x24_8steps.npz"))
    uh, st = GS.solve_burgers(meshes.rectangle(24, 24), dt=float(g["dt"]), num_steps=8, return_stats=True)
    assert rel(uh.x.array, g["uh"]) < TOL
    assert st["newton_iterations"] == int(g["newton_its"].sum())
    assert rel(st["h"], g["h"]) < 1e-11


def test_golden_kpp_reference_mesh():
    """Mesh of the reference's Data/KPP_RV.h5 (9,514 gmsh triangles)."""
    m = np.load(os.path.join(GOLD, "kpp_rv_mesh.npz"))
    g = np.load(os.path.join(GOLD, "kpp_refmesh_5steps.npz"))
    uh, st = GS.solve_kpp((m["x"], m["cells"]), dt=float(g["dt"]), num_steps=5, return_stats=True)
    assert rel(uh.x.array, g["uh"]) < TOL
    assert st["newton_iterations"] == int(g["newton_its"].sum())
    assert rel(st["eps"], g["eps"]) < 1e-8


def test_golden_advection_reference_mesh():
    """Mesh of the reference's Code/Linear_advection/Data/RV/RV_node.h5 (unit disk, h = 1/16)."""
    m = np.load(os.path.join(GOLD, "rv_node_mesh.npz"))
    g = np.load(os.path.join(GOLD, "advection_refmesh_10steps.npz"))
    uh, st = GS.solve_advection((m["x"], m["cells"]), hmax=1 / 16, num_steps=10, return_stats=True)
    assert st["dt"] == float(g["dt"]) == float(m["first_time_stamp"])   # dt formula, bit-exact
    assert rel(uh.x.array, g["uh"]) < TOL


# ---------------------------------------------------------------- BASELINE sizes, properties
@pytest.fixture(scope="module")
def big():
    n = 1024   # BASELINE.json configs[1]: 1024x1024 structured, 1,050,625 dofs
    x, c = meshes.rectangle(n, n)
    ctx = Context((x, c))
    yield n, x, c, ctx
    ctx.close()


def test_full_size_mass_and_stiffness_identities(big):
    n, x, c, ctx = big
    one = np.ones(ctx.n)
    Mone = ctx.spmv(L.MAT_MASS, one)
    assert abs(Mone.sum() - 1.0) < 1e-12                      # 1^T M 1 = |Omega| (checksum of checksums)
    hx = 1.0 / n
    interior = np.setdiff1d(np.arange(ctx.n), ctx.boundary_dofs())
    assert np.allclose(Mone[interior], hx * hx, rtol=1e-12)   # lumped mass of an interior node = 6 |K| / 3
    ctx.assemble_stiffness(None)
    assert np.abs(ctx.spmv(L.MAT_STIFFNESS, one)).max() < 1e-9
    lin = 2.0 * x[:, 0] - 3.0 * x[:, 1]
    assert np.abs(ctx.spmv(L.MAT_STIFFNESS, lin)[interior]).max() < 1e-9


def test_full_size_spmv_linearity_and_solve_roundtrip(big):
    n, x, c, ctx = big
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=ctx.n), rng.normal(size=ctx.n)
    ya, yb, yab = ctx.spmv(L.MAT_MASS, a), ctx.spmv(L.MAT_MASS, b), ctx.spmv(L.MAT_MASS, 2.0 * a - 3.0 * b)
    assert rel(yab, 2.0 * ya - 3.0 * yb) < 1e-14
    # solve -> multiply round trip: M (M^-1 y) = y
    for solver in ("chebyshev", "pcg"):
        z = ctx.solve(L.MAT_MASS, ya, solver=solver, rtol=1e-13)
        assert rel(z, a) < 1e-11 and rel(ctx.spmv(L.MAT_MASS, z), ya) < 1e-12


def test_full_size_burgers_steps(big):
    n, x, c, ctx = big
    u0 = GS.burgers_initial_condition(np.vstack([x.T, np.zeros(ctx.n)]))
    uh, st = GS.solve_burgers(ctx, dt=0.5 / n, num_steps=6, return_stats=True)
    u = uh.x.array
    assert np.all(np.isfinite(u)) and st["steps"] == 6 and st["newton_iterations"] >= 6
    assert u.min() > -1.0 - 0.05 and u.max() < 0.8 + 0.05              # RV keeps over/undershoots small
    eps = st["eps"]
    assert eps.min() >= 0.0 and np.all(eps <= 0.5 * st["h"] * np.sqrt(2.0) * 1.0 * (1 + 1e-12) * np.abs(u).max() * 1.2)
    bnd = ctx.boundary_dofs()
    from oracle.solvers import burgers_exact   # Dirichlet data honoured exactly
    assert np.array_equal(u[bnd], burgers_exact(x[bnd], st["time"]))
    # far from the fronts nothing has moved yet
    far = (np.abs(x[:, 0] - 0.5) > 0.1) & (np.abs(x[:, 1] - 0.5) > 0.1)
    assert np.abs(u[far] - u0[far]).max() < 1e-6
    # bitwise reproducible
    uh2 = GS.solve_burgers(ctx, dt=0.5 / n, num_steps=6)
    assert np.array_equal(uh2.x.array, u)


def test_kpp_4M_cells_unstructured_steps():
    """BASELINE.json configs[2]: KPP on a ~4M-cell unstructured (jittered, randomly permuted) mesh."""
    n = 1448
    x, c = meshes.jittered(n, n, (-2, -2), (2, 2))
    ctx = Context((x, c))
    assert c.shape[0] == 2 * n * n
    dt = 0.64 * 4.0 / n
    uh, st = GS.solve_kpp(ctx, dt=dt, num_steps=3, return_stats=True)
    u = uh.x.array
    assert np.all(np.isfinite(u)) and st["steps"] == 3
    assert u.min() > np.pi / 4 - 1.0 and u.max() < 3.5 * np.pi + 1.0   # oracle on coarse meshes: about -0.5 / +0.5
    assert np.array_equal(u[ctx.boundary_dofs()], np.full(ctx.boundary_dofs().size, np.pi / 4))
    # invariance under renumbering: the same mesh without the random permutation gives the same field
    x2, c2 = meshes.jittered(n, n, (-2, -2), (2, 2), permute=False)
    uh2 = GS.solve_kpp(Context((x2, c2)), dt=dt, num_steps=3)
    # match nodes by coordinates
    k1 = np.lexsort((x[:, 1], x[:, 0]))
    k2 = np.lexsort((x2[:, 1], x2[:, 0]))
    assert np.array_equal(x[k1], x2[k2])
    assert rel(u[k1], uh2.x.array[k2]) < 1e-9
    ctx.close()


def test_advection_convergence_rate_on_gpu():
    """(f-2) L2-error functional + rate fit.  Published: 2.25 / 2.11 for smooth / continuous data
    (full rotation on a disk); here a short rotation on a square, rate > 1.7."""
    from oracle.solvers import advection_initial_condition

    hs, errs = [], []
    for n in (32, 64, 128):
        x, c = meshes.rectangle(n, n, (-1, -1), (1, 1))
        ctx = Context((x, c))
        w = GS.advection_velocity(np.vstack([x.T, np.zeros(x.shape[0])])).T.copy()
        dt = GS.advection_dt(w, 2.0 / n) / 2
        steps = int(round(0.1 / dt))
        uh = GS.solve_advection(ctx, dt=dt, num_steps=steps)
        th = 2 * np.pi * dt * steps
        R = np.array([[np.cos(th), np.sin(th)], [-np.sin(th), np.cos(th)]])
        exact = advection_initial_condition(x @ R.T) * (np.linalg.norm(x, axis=1) < 0.8)
        got = uh.x.array * (np.linalg.norm(x, axis=1) < 0.8)
        hs.append(2.0 / n)
        errs.append(GS.l2_error(ctx, got, exact))
        ctx.close()
    assert GS.convergence_rate(hs, errs) > 1.7, (hs, errs)


def test_two_gpu_partition_parity():
    """Domain-decomposed run (NCCL halo exchange + all-reduce) against the oracle; needs >= 2 GPUs."""
    import subprocess
    import sys

    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run tests/dist_gpu_check.py under torchrun on a multi-GPU box)")
    script = os.path.join(os.path.dirname(__file__), "dist_gpu_check.py")
    nproc = 4 if ngpu >= 4 else 2   # 4 ranks exercise unequal ghost counts and 3 neighbours per rank
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
                        "--master-addr", "127.0.0.1", "--master-port", "29533", script],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
