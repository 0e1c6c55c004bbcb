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

PARITY UNPINNED against dolfinx itself: the reference holds no assertion, no
golden vector and no known-answer test for this path.
What *is* pinned, in ``tests/test_oracle_*.py``:

* analytic known answers on the reference's ``tests/verification`` meshes
  (``hk_test.py:36-38``, ``stiffness.py:38``, ``patch_test.py:15``);
* P1 identities (row sums of M, C·1 = 0, K·1 = 0, quadrature exactness);
* the first time stamp of ``Code/Linear_advection/Data/RV/RV_node.xdmf``
  (``dt`` formula, 16 digits) on the mesh stored in ``RV_node.h5``;
* the literal per-node Python loops of ``RV.py`` against the vectorised forms;
* observed convergence rates against the published ones.
"""
