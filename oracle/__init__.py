"""CPU oracle for the residual-viscosity (RV) P1 hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``conservation-fem_b200/`` may import
this package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may.

It is a numpy/scipy restatement of what the reference's per-time-step path
computes (reference files are cited per function, relative to the upstream
repository root):

* ``Code/Utils/RV.py:27-142``      nodal RV viscosity formulas
* ``Code/Utils/helpers.py:7-50``   nodal mesh size h_CG, smooth_vector
* ``Code/Utils/SI.py:12-28``       node patches
* ``Code/KPP/KPP_exact.py:118-166``, ``Code/Burgers_equation/Exact_Burger_RV.py:169-237``,
  ``Code/Linear_advection/RV_node_convergence.py:104-236``  the time loops

The arithmetic those scripts delegate to dolfinx 0.9.0 / FFCx 0.9.0 / basix
0.9.0 / PETSc 3.22.1 (``Environment/fenicsx-env.yml:36-43,194``; none of it is
vendored in the reference and none is installable here) is restated from the
published algorithms: closed-form P1 element integrals, the 6-point degree-4
and 7-point degree-5 symmetric triangle rules, dolfinx's Dirichlet lifting
convention and NewtonSolver stopping rules, and a sparse direct LU
(``scipy.sparse.linalg.splu``) standing in for ``KSP preonly + PC lu``.

PINNED against dolfinx itself for the linear path, UNPINNED for the nonlinear one.

The reference holds no assertion and no known-answer test for this path, but
it does store dolfinx output: three 285-frame time series on its 1,011-node unit-disk mesh,
``Code/Linear_advection/Data/RV/RV_node.h5`` (written by ``tests/eps_func.py``),
``Data/RV/RV_cell.h5`` (``RV_cell.py``) and ``Data/SI/smoothness.h5`` (the loop of
``smoothness_old_convergence.py``).  ``solvers.run_advection_stored`` restates those three scripts and
reproduces every one of the 3 x 285 stored frames to <= 1.2e-14 relative L2
(``tests/golden/make_golden.py`` checks all frames, ``tests/test_oracle_pinned.py`` the 13 committed
per series).  That pins, against dolfinx 0.9 / PETSc LU: P1 mass, convection and nodal-viscosity
stiffness assembly, the Dirichlet rows and lifting, the L2 projection of h_K, the BDF1 residual
projection and its normalisation, the pointwise and cell-based viscosities, the smoothness ratio,
every entry of the assembled Crank-Nicolson matrix (the si_old run feeds them back into alpha), and
the ``dt`` formula / time stamps.

Still unpinned against dolfinx (no stored output exists: ``Code/Burgers_equation/Data/RV/solution.xdmf``
has no ``.h5``, ``Data/KPP_RV.h5`` holds only a mesh, ``Data/RV/solution.h5`` is a P2 run): the
nonlinear flux quadrature (6-/7-point rules), the Newton loop and the patch-based ``RV.py`` formulas.
Those rest on:

* analytic known answers on the reference's ``tests/verification`` meshes
  (``hk_test.py:36-38``, ``stiffness.py:38``, ``patch_test.py:15``);
* P1 identities (row sums of M, C.1 = 0, K.1 = 0, quadrature exactness, J = dF/du);
* the literal per-node Python loops of ``RV.py`` against the vectorised forms;
* observed convergence rates against the published ones.
"""
