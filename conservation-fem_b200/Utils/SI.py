"""``Utils.SI`` — the reference ``Code/Utils/SI.py``: node patches (``SI.py:12-28``, used by the RV
solvers too) and the smoothness-indicator viscosities (``SI.py:38-67,147-192``)."""
import numpy as np

from cfem_b200 import _lib as L
from cfem_b200.context import Context
from cfem_b200.solvers import NodalFunction


class SI:
    def __init__(self, Cm, domain, eps):
        self.Cm = Cm
        self.domain = domain
        self.eps = eps
        self._ctx = domain if isinstance(domain, Context) else Context.for_domain(domain)

    def get_patch_dictionary(self):
        """node -> set of nodes sharing a cell with it, itself included."""
        return self._ctx.patch_dictionary()

    def sigmoid_activation(self, alpha):
        s, x0 = 20.0, 0.5
        return 1.0 / (1.0 + np.exp(-s * (alpha - x0)))

    def get_epsilon_nonlinear(self, velocity_field, node_patches, h_CG, u_n, stiffness_matrix=None, plot_func=None,
                              degree=1, flux=None, use_bc=True):
        """``SI.py:38-67``.  ``stiffness_matrix`` is accepted for signature parity; the unit stiffness matrix
        of the mesh (Dirichlet rows/cols as identity when ``use_bc``, as ``Exact_Burger_SI.py:169-172``
        assembles it) lives on the GPU.  ``plot_func.x.array`` receives psi(alpha) like the reference."""
        from Utils.RV import _identify_flux

        if degree != 1:
            raise NotImplementedError("the GPU path covers P1 (degree=1) only")
        eps, psi = self._ctx.si_epsilon(_identify_flux(velocity_field, flux), self.Cm, self.eps, u_n, h_CG,
                                        use_bc=use_bc, want_psi=True)
        if plot_func is not None:
            plot_func.x.array[:] = psi
        return NodalFunction(eps, "epsilon")

    def get_epsilon_linear(self, w, node_patches, h_CG, u_n, stiffness_matrix=None, numerator_func=None, degree=1,
                           use_bc=True):
        """``SI.py:147-192`` (floor 1e-8, ``||w_i||`` from the P1 velocity field)."""
        if degree != 1:
            raise NotImplementedError("the GPU path covers P1 (degree=1) only")
        eps = self._ctx.si_epsilon(L.FLUX_ADVECTION, self.Cm, 1e-8, u_n, h_CG, w=w, use_bc=use_bc)
        return NodalFunction(eps, "epsilon")
