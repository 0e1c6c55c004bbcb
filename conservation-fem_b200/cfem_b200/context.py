"""Host-side handle on one mesh: owns the C-ABI context and exposes the hot-path
operators with numpy / torch arrays in the caller's (dolfinx) numbering.

PyTorch is never required here; CUDA tensors are accepted as arguments (their
``data_ptr()`` is passed through) because the north star allows torch as the
buffer allocator.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib as L
from .meshes import as_mesh


def _field(a):
    """Accept a dolfinx-like Function (``.x.array``), numpy array or torch tensor."""
    if a is None:
        return None
    if hasattr(a, "x") and hasattr(a.x, "array"):
        a = a.x.array
    return L.f64(a)


class PatchDict(dict):
    """``SI.get_patch_dictionary`` result that also knows the GPU context it came from."""
    ctx = None


class Context:
    """Everything the reference recomputes per step but that depends only on the mesh.

    Built once from a dolfinx mesh or an ``(x, cells)`` pair: Hilbert ordering,
    node patches / CSR pattern (``Code/Utils/SI.py:12-28``), boundary dofs
    (``Code/KPP/KPP_exact.py:85-89``), mass matrices, assembly tiles.
    """

    # mesh OBJECT -> its context.  Keyed by the object itself (weakly): the entry dies with the mesh, an id() reused
    # by another object can never alias it, and the context stays alive as long as the mesh does -- a helper called
    # with nothing but the mesh (get_nodal_h(domain)) builds it once.
    _cache: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()

    def __init__(self, domain, device=0, order="hilbert", comm=None, partition=None):
        """``comm``: None (one GPU) or ``(rank, world, nccl_id_bytes)`` from ``distributed.make_comm``.
        In a distributed context every rank passes the same global mesh and global-sized fields;
        results come back at the dofs the rank owns (``distributed.allgather_field`` merges them).
        ``partition``: None = equal ranges of the Hilbert order; ``"metis"`` = METIS k-way on the nodal graph
        (computed here, identically on every rank); or an int32 array with the owning rank of every node
        (``distributed.make_partition`` computes it once and broadcasts it)."""
        lib = L.load()
        x, cells = as_mesh(domain)
        self.x = x
        self.cells = cells
        self.n = x.shape[0]
        h = C.c_void_p()
        o = L.ORDER_HILBERT if order == "hilbert" else L.ORDER_NATURAL
        self.rank, self.world = 0, 1
        if comm is None or comm[1] == 1:
            L.check(lib.cfem_create(C.byref(h), int(device), x.shape[0], cells.shape[0], L.ptr(x), 2,
                                    L.ptr(cells), 4, o))
        else:
            self.rank, self.world, nid = int(comm[0]), int(comm[1]), comm[2]
            buf = (C.c_char * 128).from_buffer_copy(bytes(nid))
            part = None
            if isinstance(partition, str):
                part = L.host_partition(x, cells, self.world, partition) if partition != "hilbert" else None
            elif partition is not None:
                part = np.ascontiguousarray(partition, dtype=np.int32)
                if part.size != x.shape[0]:
                    raise ValueError("partition must hold one rank id per node")
            self.partition = part
            L.check(lib.cfem_create_partitioned(C.byref(h), int(device), self.rank, self.world, C.addressof(buf),
                                                x.shape[0], cells.shape[0], L.ptr(x), 2, L.ptr(cells), 4, o, L.ptr(part)))
        self._h = h
        self._lib = lib
        self.n_owned = lib.cfem_num_owned(h)
        self.n_ghosts = lib.cfem_num_ghosts(h)
        self.nnz = lib.cfem_num_nonzeros(h)
        self._pattern = None
        self._h_nodal = None

    # -- life cycle
    def close(self):
        if getattr(self, "_h", None):
            self._lib.cfem_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def for_domain(cls, domain, **kw):
        """One cached context per mesh OBJECT (RV / SI / get_nodal_h share it).

        Objects that cannot be weakly referenced or hashed -- plain ``(x, cells)`` tuples, lists -- are not
        cached: every call builds a context, which the caller should then keep (``Context(domain)``)."""
        if isinstance(domain, Context):
            return domain
        try:
            ctx = cls._cache.get(domain)
        except TypeError:          # unhashable / not weak-referenceable
            return cls(domain, **kw)
        if ctx is None or ctx._h is None:
            ctx = cls(domain, **kw)
            try:
                cls._cache[domain] = ctx
            except TypeError:
                pass
        return ctx

    # -- mesh relations
    @property
    def num_tiles(self):
        return self._lib.cfem_num_tiles(self._h)

    @property
    def device_bytes(self):
        return self._lib.cfem_device_bytes(self._h)

    def csr_pattern(self):
        if self._pattern is None:
            rowptr = np.empty(self.n + 1, dtype=np.int32)
            colidx = np.empty(self.nnz, dtype=np.int32)
            L.check(self._lib.cfem_get_csr_pattern(self._h, L.ptr(rowptr), L.ptr(colidx)))
            self._pattern = (rowptr, colidx)
        return self._pattern

    def patch_dictionary(self):
        """dict node -> set(nodes) like ``SI.get_patch_dictionary``, keys in the reference's insertion order
        (first appearance of a dof when the cells are scanned in order, ``SI.py:18-26``) -- the order
        ``helpers.smooth_vector`` sweeps in.  The dict remembers its context (``.ctx``)."""
        rowptr, colidx = self.csr_pattern()
        flat = np.asarray(self.cells).reshape(-1)
        _, first = np.unique(flat, return_index=True)
        keys = flat[np.sort(first)]
        d = PatchDict((int(i), set(colidx[rowptr[i]:rowptr[i + 1]].tolist())) for i in keys)
        d.ctx = self
        return d

    def boundary_dofs(self):
        nb = self._lib.cfem_num_boundary(self._h)
        out = np.empty(nb, dtype=np.int32)
        L.check(self._lib.cfem_get_boundary_dofs(self._h, L.ptr(out)))
        return out

    def set_dirichlet(self, dofs):
        dofs = np.ascontiguousarray(dofs, dtype=np.int32)
        L.check(self._lib.cfem_set_dirichlet(self._h, L.ptr(dofs), dofs.size))

    def ordering(self):
        out = np.empty(self.n, dtype=np.int32)
        L.check(self._lib.cfem_get_ordering(self._h, L.ptr(out)))
        return out

    # -- (a-1)
    def nodal_h(self, rtol=1e-14, max_it=1000):
        if self._h_nodal is None:
            out = np.empty(self.n)
            it = C.c_int(0)
            L.check(self._lib.cfem_nodal_h(self._h, L.ptr(out), rtol, max_it, C.byref(it)))
            self._h_nodal = out
            self.nodal_h_iterations = it.value
        return self._h_nodal.copy()

    # -- (a-3)
    def rv_residual(self, flux, scheme, dt, u_n, u_old, u_oo=None, w=None, use_bc=True, R0=None,
                    rtol=1e-13, max_it=1000):
        R = np.zeros(self.n) if R0 is None else np.array(_field(R0), dtype=np.float64, copy=True)
        it = C.c_int(0)
        u_n, u_old, u_oo, w = _field(u_n), _field(u_old), _field(u_oo), _field(w)
        L.check(self._lib.cfem_rv_residual(self._h, _flux(flux), L.BDF2 if scheme in ("bdf2", 2) else L.BDF1,
                                           float(dt), L.ptr(u_n), L.ptr(u_old), L.ptr(u_oo), L.ptr(w),
                                           int(bool(use_bc)), L.ptr(R), rtol, max_it, C.byref(it)))
        self.last_iterations = it.value
        return R

    # -- (a-4..6)
    def rv_epsilon(self, variant, flux, Cvel, Crv, uh=None, u_n=None, Rh=None, h=None, w=None, out=None):
        variant = {"nonlinear": L.EPS_NONLINEAR, "linear": L.EPS_LINEAR, "pointwise": L.EPS_POINTWISE,
                   "first_order": L.EPS_FIRST_ORDER, "linear_simple": L.EPS_LINEAR_SIMPLE,
                   "cell": L.EPS_CELL}.get(variant, variant)
        eps = np.empty(self.n) if out is None else out
        uh, u_n, Rh, h, w = _field(uh), _field(u_n), _field(Rh), _field(h), _field(w)
        L.check(self._lib.cfem_rv_epsilon(self._h, variant, _flux(flux), float(Cvel), float(Crv), L.ptr(uh),
                                          L.ptr(u_n), L.ptr(Rh), L.ptr(h), L.ptr(w), L.ptr(eps)))
        return eps

    # -- (f-1) smoothness indicator
    def si_epsilon(self, flux, Cm, floor, u_n, h, w=None, use_bc=True, want_psi=False):
        eps = np.empty(self.n)
        psi = np.empty(self.n) if want_psi else None
        u_n, h, w = _field(u_n), _field(h), _field(w)
        L.check(self._lib.cfem_si_epsilon(self._h, _flux(flux), float(Cm), float(floor), int(bool(use_bc)), L.ptr(u_n),
                                          L.ptr(h), L.ptr(w), L.ptr(psi), L.ptr(eps)))
        return (eps, psi) if want_psi else eps

    # -- (a-7, a-8)
    def assemble_advection(self, dt, w, eps, u_n, bc_values=None):
        b = np.empty(self.n)
        w, eps, u_n, bc_values = _field(w), _field(eps), _field(u_n), _field(bc_values)
        L.check(self._lib.cfem_assemble_advection(self._h, float(dt), L.ptr(w), L.ptr(eps), L.ptr(u_n),
                                                  L.ptr(bc_values), L.ptr(b)))
        return b

    def assemble_cn_residual(self, flux, dt, uh, u_n, eps, bc_values=None):
        F = np.empty(self.n)
        uh, u_n, eps, bc_values = _field(uh), _field(u_n), _field(eps), _field(bc_values)
        L.check(self._lib.cfem_assemble_cn_residual(self._h, _flux(flux), float(dt), L.ptr(uh), L.ptr(u_n),
                                                    L.ptr(eps), L.ptr(bc_values), L.ptr(F)))
        return F

    def assemble_cn_jacobian(self, flux, dt, uh, eps):
        uh, eps = _field(uh), _field(eps)
        L.check(self._lib.cfem_assemble_cn_jacobian(self._h, _flux(flux), float(dt), L.ptr(uh), L.ptr(eps)))

    def assemble_stiffness(self, eps=None):
        eps = _field(eps)
        L.check(self._lib.cfem_assemble_stiffness(self._h, L.ptr(eps)))

    def matrix(self, which):
        """Context matrix as a scipy CSR in caller numbering."""
        import scipy.sparse as sp

        rowptr, colidx = self.csr_pattern()
        vals = np.empty(self.nnz)
        L.check(self._lib.cfem_matrix_values(self._h, int(which), L.ptr(vals)))
        return sp.csr_matrix((vals, colidx.copy(), rowptr.copy()), shape=(self.n, self.n))

    # -- (a-9)
    def spmv(self, which, x):
        x = _field(x)
        y = np.empty(self.n)
        L.check(self._lib.cfem_spmv(self._h, int(which), L.ptr(x), L.ptr(y)))
        return y

    def solve(self, which, b, x0=None, solver="pcg", rtol=1e-13, atol=0.0, max_it=2000):
        b = _field(b)
        x = np.zeros(self.n) if x0 is None else np.array(_field(x0), dtype=np.float64, copy=True)
        it, rr = C.c_int(0), C.c_double(0.0)
        L.check(self._lib.cfem_solve(self._h, int(which), L.SOLVER_BY_NAME.get(solver, solver), L.ptr(b), L.ptr(x),
                                     rtol, atol, max_it, C.byref(it), C.byref(rr)))
        self.last_iterations, self.last_relres = it.value, rr.value
        return x

    # -- (a-10)
    def state_set(self, uh=None, u_n=None, u_old=None, u_oo=None, RH=None, h=None, w=None, t=0.0,
                  keep_predictions=False):
        """Load state fields (None: leave as is).  ``keep_predictions`` keeps the solvers' iteration-count
        predictions from the previous call (a negative ``t`` is not meaningful, so the flag rides on a
        separate entry point: ``cfem_state_update``)."""
        args = [_field(a) for a in (uh, u_n, u_old, u_oo, RH, h, w)]
        self._keep = args
        fn = self._lib.cfem_state_update if keep_predictions else self._lib.cfem_state_set
        L.check(fn(self._h, *[L.ptr(a) for a in args], float(t)))

    def state_get(self, names=("uh",), out=None):
        """Return dict name -> array for names among uh,u_n,u_old,u_oo,RH,eps (+ 't')."""
        order = ("uh", "u_n", "u_old", "u_oo", "RH", "eps")
        bufs = {k: ((out or {}).get(k) if out and k in out else np.empty(self.n)) for k in names if k in order}
        t = C.c_double(0.0)
        L.check(self._lib.cfem_state_get(self._h, *[L.ptr(bufs.get(k)) for k in order], C.byref(t)))
        bufs["t"] = t.value
        return bufs

    def state_update_owned(self, uh=None, u_n=None, u_old=None, u_oo=None, RH=None, t=0.0):
        """``state_set(keep_predictions=True)`` with the fields in this rank's owned order (entry i is caller dof
        ``owned_dofs()[i]``): one contiguous copy per field, no global-sized arrays."""
        args = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (uh, u_n, u_old, u_oo, RH)]
        for a in args:
            if a is not None and a.size != self.n_owned:
                raise ValueError("owned-layout fields must have n_owned entries")
        self._keep = args
        L.check(self._lib.cfem_state_update_owned(self._h, *[L.ptr(a) for a in args], float(t)))

    def state_get_owned(self, names=("uh",), out=None):
        order = ("uh", "u_n", "u_old", "u_oo", "RH", "eps")
        bufs = {k: ((out or {}).get(k) if out and k in out else np.empty(self.n_owned)) for k in names if k in order}
        t = C.c_double(0.0)
        L.check(self._lib.cfem_state_get_owned(self._h, *[L.ptr(bufs.get(k)) for k in order], C.byref(t)))
        bufs["t"] = t.value
        return bufs

    def step_scalar(self, params: "L.StepParams", n_steps=1, bc_values=None):
        st = L.StepStats()
        bc_values = _field(bc_values)
        L.check(self._lib.cfem_step_scalar(self._h, C.byref(params), int(n_steps), L.ptr(bc_values), C.byref(st)))
        return st.as_dict()

    def step_scalar_si(self, params: "L.StepParams", Cm, floor=1e-8, smooth_l=0.0, smooth_order=None, n_steps=1,
                       bc_values=None):
        """Smoothness-indicator stepper (``Exact_Burger_SI.py:159-197``); ``smooth_l`` > 0 applies
        ``smooth_vector(uh, patches, smooth_l)`` after every Newton solve, sweeping in ``smooth_order``."""
        st = L.StepStats()
        bc_values = _field(bc_values)
        order = None if smooth_order is None else np.ascontiguousarray(smooth_order, dtype=np.int32)
        if order is not None and order.size != self.n:
            raise ValueError("step_scalar_si: smooth_order must list every dof once")
        L.check(self._lib.cfem_step_scalar_si(self._h, C.byref(params), float(Cm), float(floor), float(smooth_l),
                                              L.ptr(order), int(n_steps), L.ptr(bc_values), C.byref(st)))
        return st.as_dict()

    def smooth_vector(self, u, l, order=None):
        """``helpers.smooth_vector`` in place on ``u`` (numpy array or dolfinx-like Function); ``order``: caller
        dof ids in sweep order (None = ascending)."""
        a = _field(u)
        order = None if order is None else np.ascontiguousarray(order, dtype=np.int32)
        if order is not None and order.size != self.n:
            raise ValueError("smooth_vector: order must list every dof once")
        L.check(self._lib.cfem_smooth_vector(self._h, L.ptr(a), L.ptr(order), float(l)))
        return a

    def step_advection(self, params: "L.StepParams", n_steps=1, first_gfem=False):
        st = L.StepStats()
        L.check(self._lib.cfem_step_advection(self._h, C.byref(params), int(n_steps), int(bool(first_gfem)),
                                              C.byref(st)))
        return st.as_dict()

    # -- (a-12) Euler system: (Nn,4) arrays
    def euler_state_set(self, Uh=None, Un=None, Uold=None, Uoo=None, bc_state=None, h=None, t=0.0):
        args = [None if a is None else np.ascontiguousarray(a, dtype=np.float64).reshape(self.n, 4)
                for a in (Uh, Un, Uold, Uoo, bc_state)]
        hh = _field(h)
        L.check(self._lib.cfem_euler_state_set(self._h, *[L.ptr(a) for a in args], L.ptr(hh), float(t)))

    def euler_state_get(self, want=("Uh",), out=None):
        """``out``: optional dict of caller arrays to fill (e.g. pinned buffers), keyed like ``want``."""
        given = out or {}
        out = {}

        def buf(key, shape):
            if key not in want:
                return None
            a = given.get(key)
            if a is None:
                return np.zeros(shape)
            if a.dtype != np.float64 or not a.flags.c_contiguous or a.shape != shape:
                raise ValueError(f"euler_state_get: out[{key!r}] must be a C-contiguous float64 array of shape {shape}")
            return a

        Uh, R, eps = buf("Uh", (self.n, 4)), buf("R", (self.n, 4)), buf("eps", (self.n,))
        t = C.c_double(0.0)
        L.check(self._lib.cfem_euler_state_get(self._h, L.ptr(Uh), L.ptr(R), L.ptr(eps), C.byref(t)))
        for k, v in (("Uh", Uh), ("R", R), ("eps", eps)):
            if v is not None:
                out[k] = v
        out["t"] = t.value
        return out

    def step_euler(self, params: "L.StepParams", n_steps=1):
        st = L.StepStats()
        L.check(self._lib.cfem_step_euler(self._h, C.byref(params), int(n_steps), C.byref(st)))
        return st.as_dict()

    def time_kernel(self, kernel, flux, reps=20):
        ms, by = C.c_double(0.0), C.c_double(0.0)
        L.check(self._lib.cfem_time_kernel(self._h, int(kernel), _flux(flux), int(reps), C.byref(ms), C.byref(by)))
        return ms.value, by.value

    PROFILE_CATEGORIES = ("spmv", "asm_vector", "asm_matrix", "krylov_vector", "rv", "misc", "chebyshev", "comm", "solver")

    def profile_begin(self, max_launches=200000):
        L.check(self._lib.cfem_profile_begin(self._h, int(max_launches)))

    def profile_gaps(self):
        """Stream idle time (ms) after the scopes of each category during the profiled calls; before profile_end."""
        ms = (C.c_double * len(self.PROFILE_CATEGORIES))()
        L.check(self._lib.cfem_profile_gaps(self._h, ms))
        return {k: ms[i] for i, k in enumerate(self.PROFILE_CATEGORIES)}

    def profile_end(self):
        ms = (C.c_double * len(self.PROFILE_CATEGORIES))()
        cnt = (C.c_int64 * len(self.PROFILE_CATEGORIES))()
        L.check(self._lib.cfem_profile_end(self._h, ms, cnt))
        return {k: {"ms": ms[i], "launches": cnt[i]} for i, k in enumerate(self.PROFILE_CATEGORIES)}

    def owned_dofs(self):
        """Caller-numbered dofs this rank owns (all of them on one GPU)."""
        return self.ordering()[: self.n_owned]

    def comm_stats(self):
        a, b, d = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        L.check(self._lib.cfem_comm_stats(self._h, C.byref(a), C.byref(b), C.byref(d)))
        return {"halo_exchanges": a.value, "allreduces": b.value, "halo_doubles_sent_per_exchange": d.value}

    # -- (f-2) L2 error against the P3 interpolant of an exact solution
    P3_NODES = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 2 / 3, 1 / 3], [0, 1 / 3, 2 / 3], [2 / 3, 0, 1 / 3],
                         [1 / 3, 0, 2 / 3], [2 / 3, 1 / 3, 0], [1 / 3, 2 / 3, 0], [1 / 3, 1 / 3, 1 / 3]])

    def p3_cell_points(self):
        """(Nc, 10, 2) coordinates of the P3 Lagrange nodes of every cell, in the node order ``cfem_l2_error_p3``
        expects: the cell's 3 vertices, 2 per edge (opposite vertex 0, 1, 2), centroid."""
        return np.einsum("pk,ckd->cpd", self.P3_NODES, self.x[self.cells][:, :, :2])

    def l2_error_p3(self, uh, exact):
        """``sqrt(assemble_scalar((uh - u_exact)**2 * dx))`` with ``u_exact = Function(P3).interpolate(exact)``
        (``Exact_Burger_RV_conv.py:81-86,223``).  ``exact``: callable on ``x`` with shape ``(3, N)`` like
        ``Function.interpolate``, or a ready ``(Nc, 10)`` table at ``p3_cell_points()``.  ``uh`` None: the resident uh."""
        if callable(exact):
            pts = self.p3_cell_points().reshape(-1, 2)
            X = np.zeros((3, pts.shape[0]))
            X[0], X[1] = pts[:, 0], pts[:, 1]
            table = np.asarray(exact(X), dtype=np.float64).reshape(-1, 10)
        else:
            table = np.asarray(exact, dtype=np.float64).reshape(-1, 10)
        if table.shape[0] != self.cells.shape[0]:
            raise ValueError("exact-solution table must have one row of 10 values per cell")
        table = np.ascontiguousarray(table)
        err = C.c_double(0.0)
        L.check(self._lib.cfem_l2_error_p3(self._h, L.ptr(_field(uh)), L.ptr(table), C.byref(err)))
        return err.value

    def comm_timers(self, reset=True):
        """Device-measured waits of the peer-memory data plane since the last reset (microseconds / counts)."""
        out = (C.c_double * 12)()
        L.check(self._lib.cfem_comm_timers(self._h, out, int(bool(reset))))
        k = ("halo_wait_us_total", "halo_waits", "halo_wait_us_max", "allreduce_us_total", "allreduces",
             "allreduce_us_max", "barrier_us_worker0", "barriers", "tile_halo_wait_us_total", "tile_halo_waits",
             "tile_halo_wait_us_max", "reserved")
        return dict(zip(k, [float(v) for v in out]))

    def synchronize(self):
        L.check(self._lib.cfem_synchronize(self._h))


def _flux(flux):
    if isinstance(flux, str):
        return L.FLUX_BY_NAME[flux]
    return int(flux)


def step_params(flux, dt, Cvel, Crv, scheme="bdf2", newton_rtol=1e-4, newton_atol=1e-10, newton_max_it=100,
                solver="bicgstab", lin_rtol=1e-13, lin_max_it=2000, bc_kind="constant", bc_value=0.0,
                residual_bc=True, mass_solver="chebyshev", mass_rtol=0.0):
    p = L.StepParams()
    p.flux = _flux(flux)
    p.scheme = L.BDF2 if scheme in ("bdf2", 2) else L.BDF1
    p.dt, p.Cvel, p.Crv = float(dt), float(Cvel), float(Crv)
    p.newton_rtol, p.newton_atol, p.newton_max_it = newton_rtol, newton_atol, newton_max_it
    p.solver = L.SOLVER_BY_NAME.get(solver, solver)
    p.lin_rtol, p.lin_max_it = lin_rtol, lin_max_it
    p.bc_kind = {"constant": L.BC_CONSTANT, "burgers_exact": L.BC_BURGERS_EXACT, "user": L.BC_USER}.get(bc_kind, bc_kind)
    p.bc_value = float(bc_value)
    p.residual_bc = int(bool(residual_bc))
    p.mass_solver = 100 + L.SOLVER_PCG if mass_solver == "pcg" else 0
    p.mass_rtol = float(mass_rtol)
    return p
