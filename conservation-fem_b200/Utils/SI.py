"""``Utils.SI`` — only the part of the reference ``Code/Utils/SI.py`` that the RV
solvers use: ``SI(Cm, domain, eps).get_patch_dictionary()`` (``SI.py:7-28``).
The smoothness-indicator viscosities are outside the RV hot path."""
from cfem_b200.context import Context


class SI:
    def __init__(self, Cm, domain, eps):
        self.Cm = Cm
        self.domain = domain
        self.eps = eps
        self._ctx = domain if isinstance(domain, Context) else Context.for_domain(domain)

    def get_patch_dictionary(self):
        """node -> set of nodes sharing a cell with it, itself included."""
        return self._ctx.patch_dictionary()
