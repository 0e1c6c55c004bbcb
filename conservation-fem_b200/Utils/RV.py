"""``Utils.RV`` — residual-viscosity class, API of the reference ``Code/Utils/RV.py``.

Every method keeps the reference signature and returns a new function object
(``.x.array``); ``velocity_field`` stays a callable ``u -> f'(u)`` (nonlinear
variants) or a P1 vector function (linear variants).  The callable is only probed
at two sample values to recognise the flux (Burgers ``(u,u)`` or KPP
``(cos u, -sin u)``); pass ``flux="burgers"|"kpp"`` to skip the probe.
"""
import numpy as np

from cfem_b200 import _lib as L
from cfem_b200.context import Context
from cfem_b200.solvers import NodalFunction


def _identify_flux(velocity_field, flux=None):
    if flux is not None:
        return L.FLUX_BY_NAME[flux] if isinstance(flux, str) else int(flux)
    kind = getattr(velocity_field, "flux_kind", None)
    if kind is not None:
        return L.FLUX_BY_NAME[kind]
    probes = (0.3, -1.1)
    vals = [np.array(velocity_field(p), dtype="float").ravel() for p in probes]
    if all(np.allclose(v, [p, p]) for v, p in zip(vals, probes)):
        return L.FLUX_BURGERS
    if all(np.allclose(v, [np.cos(p), -np.sin(p)]) for v, p in zip(vals, probes)):
        return L.FLUX_KPP
    raise ValueError("velocity_field is neither the Burgers nor the KPP flux derivative; "
                     "no GPU kernel exists for it (there is no CPU fallback)")


class RV:
    def __init__(self, Cvel, Crv, domain):
        self.Cvel = Cvel
        self.Crv = Crv
        self.domain = domain
        self._ctx = domain if isinstance(domain, Context) else Context.for_domain(domain)

    def _check_degree(self, degree):
        if degree != 1:
            raise NotImplementedError("the GPU path covers P1 (degree=1) only")

    def get_epsilon(self, uh, velocity_field, residual, h, degree=1, flux=None):
        """``RV.py:27-40``: min(Cvel h |f'(u)|, Crv h^2 |R|), pointwise."""
        self._check_degree(degree)
        eps = self._ctx.rv_epsilon("pointwise", _identify_flux(velocity_field, flux), self.Cvel, self.Crv,
                                   uh=uh, Rh=residual, h=h)
        return NodalFunction(eps, "epsilon")

    def get_epsilon_1storder(self, uh, velocity_field, residual, h, degree=1, flux=None):
        """``RV.py:42-54``: 0.5 h |f'(u)|."""
        self._check_degree(degree)
        eps = self._ctx.rv_epsilon("first_order", _identify_flux(velocity_field, flux), self.Cvel, self.Crv,
                                   uh=uh, h=h)
        return NodalFunction(eps, "epsilon")

    def get_epsilon_nonlinear(self, uh, u_n, velocity_field, Rh, h_CG, node_patches=None, degree=1, flux=None):
        """``RV.py:56-90``.  ``node_patches`` is accepted for signature parity; the patches
        are the mesh's own P1 graph, already resident on the GPU."""
        self._check_degree(degree)
        eps = self._ctx.rv_epsilon("nonlinear", _identify_flux(velocity_field, flux), self.Cvel, self.Crv,
                                   uh=uh, u_n=u_n, Rh=Rh, h=h_CG)
        return NodalFunction(eps, "epsilon")

    def get_epsilon_linear(self, uh, u_n, velocity_field, Rh, h_CG, node_patches=None, degree=1):
        """``RV.py:92-127``; ``velocity_field`` is the P1 velocity function w."""
        self._check_degree(degree)
        eps = self._ctx.rv_epsilon("linear", L.FLUX_ADVECTION, self.Cvel, self.Crv, uh=uh, u_n=u_n, Rh=Rh,
                                   h=h_CG, w=velocity_field)
        return NodalFunction(eps, "epsilon")

    def get_epsilon_linear_simple(self, w, residual, u_n, h, degree=1):
        """``RV.py:129-142``; like the reference, ``residual`` is normalised IN PLACE."""
        self._check_degree(degree)
        r = residual.x.array if hasattr(residual, "x") else residual
        if not (isinstance(r, np.ndarray) and r.dtype == np.float64 and r.flags["C_CONTIGUOUS"]):
            raise TypeError("residual must be a contiguous float64 array (it is modified in place)")
        eps = self._ctx.rv_epsilon("linear_simple", L.FLUX_ADVECTION, self.Cvel, self.Crv, u_n=u_n, Rh=r, h=h, w=w)
        return NodalFunction(eps, "epsilon")
