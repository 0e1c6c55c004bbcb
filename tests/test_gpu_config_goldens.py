"""GPU vs the oracle ON the BASELINE.json configurations (not on reduced cases).

The fixtures come from ``tests/golden/make_config_goldens.py``: the oracle (numpy + SuperLU) run offline on
  configs[0]  linear advection, unit square 100x100, 50 steps               -> full field
  configs[1]  Burgers, 1024x1024 structured, 3 steps (Newton counts 1,2,1)  -> every 97th node + norms
  configs[2]  KPP, 1448x1448 jittered + randomly renumbered (4.19 M cells), 2 steps -> every 193rd node + norms
Tolerance: the north star's 1e-10 relative L2 on the field; Newton iteration counts must be identical.
All calls go through the C ABI (``cfem_b200.solvers`` -> ctypes -> libcfem_b200.so).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from cfem_b200 import meshes  # noqa: E402
from cfem_b200 import solvers as GS  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-10
# solver tolerances (on the row-equilibrated residual): the library default, and what bench.py times with
SETTINGS = [pytest.param(dict(lin_rtol=1e-13), id="default_1e-13"),
            pytest.param(dict(lin_rtol=1e-11, mass_rtol=1e-11), id="bench_1e-11")]


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


def _load(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated (tests/golden/make_config_goldens.py)")
    return np.load(path)


def test_config0_advection_100x100_50_steps():
    g = _load("config0_advection_100x100_50steps.npz")
    n = int(g["n"])
    x, c = meshes.rectangle(n, n)
    uh, st = GS.solve_advection((x, c), hmax=1.0 / n, num_steps=int(g["steps"]), return_stats=True)
    assert st["dt"] == float(g["dt"])                      # RV_node.py:82-86, bit-exact
    assert rel(st["h"], g["h"]) < 1e-12
    assert rel(uh.x.array, g["uh"]) < TOL
    assert rel(st["eps"], g["eps"]) < 1e-8


@pytest.mark.parametrize("tols", SETTINGS)
def test_config1_burgers_1024x1024_3_steps(tols):
    g = _load("config1_burgers_1024x1024_3steps.npz")
    n = int(g["n"])
    x, c = meshes.rectangle(n, n)
    uh, st = GS.solve_burgers((x, c), dt=float(g["dt"]), num_steps=int(g["steps"]), return_stats=True, **tols)
    idx = g["index"]
    u = uh.x.array
    assert rel(u[idx], g["uh"]) < TOL
    assert abs(np.linalg.norm(u) - float(g["norm"])) < TOL * float(g["norm"])
    assert st["newton_iterations"] == int(g["newton_its"].sum())
    assert rel(st["h"][idx], g["h"]) < 1e-12
    assert rel(st["RH"][idx], g["RH"]) < 1e-9 and abs(np.linalg.norm(st["RH"]) / float(g["norm_RH"]) - 1) < 1e-9
    assert rel(st["eps"][idx], g["eps"]) < 1e-8 and abs(np.linalg.norm(st["eps"]) / float(g["norm_eps"]) - 1) < 1e-8


@pytest.mark.parametrize("tols", SETTINGS)
def test_config2_kpp_4M_cells_2_steps(tols):
    g = _load("config2_kpp_1448x1448_jittered_2steps.npz")
    n = int(g["n"])
    x, c = meshes.jittered(n, n, (-2.0, -2.0), (2.0, 2.0))
    uh, st = GS.solve_kpp((x, c), dt=float(g["dt"]), num_steps=int(g["steps"]), return_stats=True, **tols)
    idx = g["index"]
    u = uh.x.array
    assert rel(u[idx], g["uh"]) < TOL
    assert abs(np.linalg.norm(u) - float(g["norm"])) < TOL * float(g["norm"])
    assert st["newton_iterations"] == int(g["newton_its"].sum())
    assert rel(st["RH"][idx], g["RH"]) < 1e-9
    assert rel(st["eps"][idx], g["eps"]) < 1e-8
