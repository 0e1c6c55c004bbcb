"""``Utils.helpers`` — ``get_nodal_h`` of the reference ``Code/Utils/helpers.py:7-38``."""
from cfem_b200.context import Context
from cfem_b200.solvers import NodalFunction


def get_nodal_h(domain, degree=1):
    """Nodal mesh size: L2 projection onto P1 of the per-cell shortest edge."""
    if degree != 1:
        raise NotImplementedError("the GPU path covers P1 (degree=1) only")
    ctx = domain if isinstance(domain, Context) else Context.for_domain(domain)
    return NodalFunction(ctx.nodal_h(), "h_CG")
