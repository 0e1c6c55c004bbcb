"""CPU: the oracle against dolfinx itself.

The reference stores three 285-frame dolfinx time series on its 1,011-node unit-disk mesh
(``Code/Linear_advection/Data/RV/RV_node.h5``, ``.../RV/RV_cell.h5``, ``.../SI/smoothness.h5``).
``tests/golden/make_golden.py`` lifts 13 frames of each into ``tests/golden/ref_series_*.npz`` (and
checks all 285 when it runs); here the oracle's restatement of the scripts that wrote them
(``oracle.solvers.run_advection_stored``) must land on the stored frames.  Everything the linear
advection hot path does is on this line: P1 mass / convection / nodal-viscosity stiffness assembly,
dolfinx's Dirichlet rows + lifting, the LU solves, the L2 projection of h_K, the BDF1 residual
projection, its normalisation, the pointwise and cell-based viscosities, the smoothness-indicator
ratio and (si_old) every entry of the assembled Crank-Nicolson matrix.

Tolerance: 1e-12 relative L2 per frame (observed 1e-14 after 285 steps; LU round-off only).
"""
import os

import numpy as np
import pytest

from oracle import solvers as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-12


@pytest.fixture(scope="module")
def disk():
    d = np.load(os.path.join(GOLD, "rv_node_mesh.npz"))
    return d["x"], d["cells"]


@pytest.mark.parametrize("variant", ["eps_func", "rv_cell", "si_old"])
def test_oracle_reproduces_stored_dolfinx_series(disk, variant):
    x, c = disk
    g = np.load(os.path.join(GOLD, f"ref_series_{variant}.npz"))
    U, m, dt = S.run_advection_stored(x, c, variant)
    assert U.shape[0] == 285 == int(np.ceil(1.0 / dt))
    # time stamps of the XDMF file: t_k = (k + 1) dt accumulated the way the scripts do (t += dt)
    t, acc = 0.0, []
    for _ in range(285):
        t += dt
        acc.append(t)
    assert np.array_equal(np.array(acc)[g["index"]], g["times"])
    for k, F in zip(g["index"], g["frames"]):
        err = np.linalg.norm(U[k] - F) / np.linalg.norm(F)
        assert err < TOL, (variant, int(k), err)


def test_stored_series_are_not_interchangeable(disk):
    """The three stored runs differ from each other by far more than the tolerance from the second
    frame on, so matching one of them is a statement about that viscosity formula."""
    g = {v: np.load(os.path.join(GOLD, f"ref_series_{v}.npz"))["frames"] for v in ("eps_func", "rv_cell", "si_old")}
    assert np.array_equal(g["eps_func"][0], g["rv_cell"][0]) and np.array_equal(g["eps_func"][0], g["si_old"][0])
    for a, b in (("eps_func", "rv_cell"), ("eps_func", "si_old"), ("rv_cell", "si_old")):
        d = np.linalg.norm(g[a][1] - g[b][1]) / np.linalg.norm(g[a][1])
        assert d > 1e-6, (a, b, d)
