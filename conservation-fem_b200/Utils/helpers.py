"""``Utils.helpers`` — ``get_nodal_h`` and ``smooth_vector`` of the reference ``Code/Utils/helpers.py:7-50``."""
import numpy as np

from cfem_b200.context import Context
from cfem_b200.solvers import NodalFunction


def get_nodal_h(domain, degree=1):
    """Nodal mesh size: L2 projection onto P1 of the per-cell shortest edge."""
    if degree != 1:
        raise NotImplementedError("the GPU path covers P1 (degree=1) only")
    ctx = domain if isinstance(domain, Context) else Context.for_domain(domain)
    return NodalFunction(ctx.nodal_h(), "h_CG")


def smooth_vector(u, patches, l):
    """``helpers.py:40-50``: in-place sweep ``u_i <- (sum_{j != i} u_j + (l-1) d_i u_i) / (l d_i)`` over the nodes
    in the key order of ``patches``, every node seeing the already smoothed values of earlier ones.  Runs on the
    GPU (level-scheduled, ``cfem_smooth_vector``).  ``patches`` must come from ``SI.get_patch_dictionary`` (it
    carries the context) or ``u`` must be a dolfinx Function (its mesh identifies the context)."""
    ctx = getattr(patches, "ctx", None)
    if ctx is None:
        ctx = Context.for_domain(u.function_space.mesh)
    ctx.smooth_vector(u, l, order=np.fromiter(patches.keys(), dtype=np.int32, count=len(patches)))
