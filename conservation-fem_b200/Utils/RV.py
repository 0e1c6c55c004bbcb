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

    def _check_patches(self, node_patches):
        """The kernels use the mesh's own P1 graph.  A ``node_patches`` argument that did not come from this
        context (``SI.get_patch_dictionary``) must describe the same graph -- anything else cannot be honoured
        and is refused instead of being silently ignored."""
        if node_patches is None or getattr(node_patches, "ctx", None) is self._ctx:
            return
        rowptr, colidx = self._ctx.csr_pattern()
        if len(node_patches) != self._ctx.n:
            raise ValueError("node_patches does not belong to this mesh (size mismatch)")
        rng = np.random.default_rng(0)
        for i in rng.integers(0, self._ctx.n, size=min(64, self._ctx.n)):
            if set(int(j) for j in node_patches[int(i)]) != set(colidx[rowptr[i]:rowptr[i + 1]].tolist()):
                raise ValueError("node_patches differs from the P1 graph of the mesh; custom patches are not supported")

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
        """``RV.py:56-90``.  The patches are the mesh's own P1 graph, already resident on the GPU;
        ``node_patches`` is checked against it (see ``_check_patches``)."""
        self._check_degree(degree)
        self._check_patches(node_patches)
        eps = self._ctx.rv_epsilon("nonlinear", _identify_flux(velocity_field, flux), self.Cvel, self.Crv,
                                   uh=uh, u_n=u_n, Rh=Rh, h=h_CG)
        return NodalFunction(eps, "epsilon")

    def get_epsilon_linear(self, uh, u_n, velocity_field, Rh, h_CG, node_patches=None, degree=1):
        """``RV.py:92-127``; ``velocity_field`` is the P1 velocity function w."""
        self._check_degree(degree)
        self._check_patches(node_patches)
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
