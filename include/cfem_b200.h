/* cfem_b200 — C ABI of the B200-native residual-viscosity (RV) P1 hot path.
 *
 * The reference (alleswe2k/Conservation-FEM) has no FFI of its own: its hot
 * path is Python calling dolfinx/PETSc.  Each entry point below therefore names
 * the reference call site (file:line, relative to the upstream repo root) whose
 * arithmetic it replaces.  INTEGRATION.md shows the ctypes stubs a maintainer
 * adds to Code/Utils to bind them.
 *
 * Conventions
 *  - Plain C: pointers, sizes, scalars.  No torch / C++ types.
 *  - Every array argument is in the CALLER's node numbering (dolfinx dof ==
 *    geometry-node index for P1) and may live in host OR device memory; the
 *    library detects which (cudaPointerGetAttributes) and stages host buffers
 *    itself.  Internally nodes and cells are re-ordered along a Hilbert curve;
 *    that permutation never leaks.
 *  - All real data are fp64, indices int32 (dolfinx 0.9 dofmaps, PetscInt=4).
 *  - Return value: 0 = ok, negative = error; cfem_last_error() has the text.
 *  - No CPU fallback: every compute entry point needs a CUDA device.
 *  - Thread model: one context per host thread / per GPU; calls are
 *    synchronous on return unless stated.
 */
#ifndef CFEM_B200_H
#define CFEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cfem_ctx cfem_ctx;

/* ---- enums (ints across the ABI) ------------------------------------- */
enum { CFEM_FLUX_ADVECTION = 0, CFEM_FLUX_BURGERS = 1, CFEM_FLUX_KPP = 2 };
enum { CFEM_BDF1 = 1, CFEM_BDF2 = 2 };
enum { CFEM_EPS_NONLINEAR = 0,     /* RV.get_epsilon_nonlinear   RV.py:56-90   */
       CFEM_EPS_LINEAR = 1,        /* RV.get_epsilon_linear      RV.py:92-127  */
       CFEM_EPS_POINTWISE = 2,     /* RV.get_epsilon             RV.py:27-40   */
       CFEM_EPS_FIRST_ORDER = 3,   /* RV.get_epsilon_1storder    RV.py:42-54   */
       CFEM_EPS_LINEAR_SIMPLE = 4, /* RV.get_epsilon_linear_simple RV.py:129-142 */
       CFEM_EPS_CELL = 5           /* per-cell loop of Code/Linear_advection/RV_cell.py:174-192 (needs u_n, Rh, w; h unused) */ };
enum { CFEM_MAT_MASS = 0,          /* int u v, no Dirichlet rows              */
       CFEM_MAT_MASS_BC = 1,       /* same, Dirichlet rows/cols -> identity   */
       CFEM_MAT_SYSTEM = 2,        /* last assembled CN matrix / Jacobian     */
       CFEM_MAT_STIFFNESS = 3      /* last assembled int eps grad u . grad v  */ };
enum { CFEM_SOLVER_PCG = 0, CFEM_SOLVER_BICGSTAB = 1, CFEM_SOLVER_GMRES = 2,
       CFEM_SOLVER_CHEBYSHEV = 3 /* mass matrices only: spectrum of D^-1 M in [1/2,2] */ };
enum { CFEM_BC_CONSTANT = 0,       /* g = bc_value             KPP_exact.py:88 */
       CFEM_BC_BURGERS_EXACT = 1,  /* g = exact Riemann soln   Exact_Burger_RV.py:37-66,172-176 */
       CFEM_BC_USER = 2            /* g given per call, one value per Dirichlet dof */ };
enum { CFEM_ORDER_HILBERT = 0, CFEM_ORDER_NATURAL = 1 };

const char* cfem_last_error(void);
int cfem_version(void);
/* sizeof(cfem_step_params) (which = 0) / sizeof(cfem_step_stats) (which = 1): lets a binding check its struct layout */
int cfem_struct_size(int which);
/* number of CUDA devices visible (0 if none / no driver) */
int cfem_device_count(void);

/* ---- context ---------------------------------------------------------
 * Builds, once per mesh, everything the per-step path re-uses: Hilbert
 * ordering, vertex->cell adjacency, the P1 CSR pattern (== the node patches of
 * SI.get_patch_dictionary, Code/Utils/SI.py:12-28), assembly tiles, boundary
 * dofs (mesh.locate_entities_boundary(all) + locate_dofs_topological,
 * Code/KPP/KPP_exact.py:85-89), the mass matrices and their Jacobi diagonals.
 *   x      host, n_nodes*xdim doubles (domain.geometry.x, xdim 2 or 3; z ignored)
 *   cells  host, n_cells*3 indices of cell_index_bytes (4 or 8) each
 *          (geometry.dofmap == V.dofmap.list for P1)
 *   order  CFEM_ORDER_*
 * All boundary dofs are Dirichlet dofs by default (as in every reference loop). */
int cfem_create(cfem_ctx** out, int device, int64_t n_nodes, int64_t n_cells,
                const double* x, int xdim, const void* cells, int cell_index_bytes,
                int order);
/* ---- multi-GPU: one process per GPU, one context per process ----
 * Every rank passes the SAME global mesh; the library orders it along the Hilbert curve,
 * gives rank r the r-th contiguous range of nodes (its rows) plus one layer of ghost nodes and
 * all cells touching an owned node, so assembly needs no reverse exchange.  Ghost values are
 * refreshed by grouped ncclSend/ncclRecv before every SpMV / assembly, Krylov dot products and
 * the RV normalisation (sum/min/max) by ncclAllReduce.  nccl_id128: the 128-byte ncclUniqueId
 * made by cfem_nccl_unique_id on rank 0 and broadcast by the caller (torch.distributed).
 * Field arguments of every other entry point stay GLOBAL-sized arrays in caller numbering;
 * outputs are written at the dofs this rank owns.  Calls are collective. */
int cfem_nccl_unique_id(void* out128);
int cfem_create_distributed(cfem_ctx** out, int device, int rank, int world, const void* nccl_id128,
                            int64_t n_nodes, int64_t n_cells, const double* x, int xdim,
                            const void* cells, int cell_index_bytes, int order);
/* Same with the partition given by the caller: node_part[i] in [0, world) is the rank that owns caller node i
 * (every rank passes the same array), e.g. from cfem_host_partition.  A part becomes a contiguous range of the
 * internal order (nodes sorted by part, then along the Hilbert curve); node_part == NULL is cfem_create_distributed. */
int cfem_create_partitioned(cfem_ctx** out, int device, int rank, int world, const void* nccl_unique_id_128,
                            int64_t n_nodes, int64_t n_cells, const double* x, int xdim, const void* cells,
                            int cell_index_bytes, int order, const int32_t* node_part);
/* Graph partition of the mesh, host only: method CFEM_PART_METIS = METIS k-way on the nodal graph (the partitioner family
 * dolfinx uses for the reference's MPI runs, Environment/fenicsx-env.yml:171,192); CFEM_PART_HILBERT = equal ranges
 * of the Hilbert order (what node_part == NULL gives).  part_out: n_nodes entries, caller numbering. */
enum { CFEM_PART_HILBERT = 0, CFEM_PART_METIS = 1 };
int cfem_host_partition(int method, int world, int64_t n_nodes, int64_t n_cells, const double* x, int xdim,
                        const void* cells, int cell_index_bytes, int32_t* part_out);
int64_t cfem_num_owned(const cfem_ctx* ctx);
int64_t cfem_num_ghosts(const cfem_ctx* ctx);
int cfem_comm_stats(const cfem_ctx* ctx, int64_t* halo_exchanges, int64_t* allreduces, int64_t* halo_doubles_sent);
/* Wait accounting of the peer-memory data plane since the last reset, measured on the device (no profiler):
 * out[0..2] halo waits: total us over all waiting CTAs, number of waits, longest single wait (us);
 * out[3..5] the cross-rank part of in-kernel all-reduces: total us, count, longest;
 * out[6..7] time the first worker CTA of the persistent solver spent in grid barriers: total us, count;
 * out[8..10] halo-word polls of the staged tile kernels (Chebyshev chain, stand-alone SpMV): per boundary tile the
 *            wait of the thread that polls the tile's last ghost column: total us, count, longest (us). */
int cfem_comm_timers(cfem_ctx* ctx, double out[12], int reset);
void cfem_destroy(cfem_ctx* ctx);
int cfem_synchronize(cfem_ctx* ctx);

int64_t cfem_num_nodes(const cfem_ctx* ctx);
int64_t cfem_num_cells(const cfem_ctx* ctx);
int64_t cfem_num_nonzeros(const cfem_ctx* ctx);
int64_t cfem_num_boundary(const cfem_ctx* ctx);
int64_t cfem_num_dirichlet(const cfem_ctx* ctx);
int64_t cfem_num_tiles(const cfem_ctx* ctx);
/* bytes of device memory the context holds */
int64_t cfem_device_bytes(const cfem_ctx* ctx);
/* L2 facts of `device`: out[0] L2 bytes, out[1] largest persisting set-aside, out[2] largest access-policy window,
 * out[3] SM count.  (The solvers keep the matrix of a solve in persisting L2 lines; CFEM_L2PERSIST=0 turns that off.) */
int cfem_device_limits(int device, int64_t out[4]);

/* Node patches / sparsity pattern in caller numbering, columns ascending.
 * rowptr: n_nodes+1, colidx: nnz (host).          SI.py:12-28 */
int cfem_get_csr_pattern(cfem_ctx* ctx, int32_t* rowptr, int32_t* colidx);
/* Sorted boundary dofs (host, cfem_num_boundary entries).  KPP_exact.py:85-89 */
int cfem_get_boundary_dofs(cfem_ctx* ctx, int32_t* dofs);
/* Replace the Dirichlet set (host array, caller numbering); re-assembles MASS_BC. */
int cfem_set_dirichlet(cfem_ctx* ctx, const int32_t* dofs, int64_t n);
/* internal->caller permutation (host, n_nodes) — exposed for tests only */
int cfem_get_ordering(cfem_ctx* ctx, int32_t* internal_to_user);

/* ---- (a-1) helpers.get_nodal_h   Code/Utils/helpers.py:7-38 ------------
 * h_K = min edge, b_i = sum h_K|K|/3, Jacobi-PCG on M h = b. */
int cfem_nodal_h(cfem_ctx* ctx, double* h_out, double rtol, int max_it, int* iters);

/* ---- (a-3) RV residual projection -------------------------------------
 * M R = int [D_t u + f'(u_n).grad u_n] phi ;  KPP_exact.py:123-137,
 * Exact_Burger_RV.py:187-203 (BDF2), Exact_Burger_RV_conv.py:186 (BDF1),
 * RV_node.py:209-214 (advection, w = P1 velocity, interleaved (Nn,2)).
 * use_bc != 0: R = 0 on Dirichlet dofs (bc0).  R_io holds the initial guess
 * on entry (warm start) and the result on exit. */
int cfem_rv_residual(cfem_ctx* ctx, int flux, int scheme, double dt,
                     const double* u_n, const double* u_old, const double* u_oo,
                     const double* w, int use_bc, double* R_io,
                     double rtol, int max_it, int* iters);

/* ---- (a-4..a-6) nodal RV viscosity   Code/Utils/RV.py:27-142 -----------
 * variant CFEM_EPS_*; flux selects ||f'(u)||_2 (Burgers sqrt(2u^2), KPP
 * sqrt(cos^2+sin^2), advection ||w_i||).  Unused inputs may be NULL.
 * LINEAR_SIMPLE also overwrites Rh with the normalised residual (RV.py:132). */
int cfem_rv_epsilon(cfem_ctx* ctx, int variant, int flux, double Cvel, double Crv,
                    const double* uh, const double* u_n, double* Rh,
                    const double* h, const double* w, double* eps_out);

/* ---- (f-1) smoothness-indicator viscosity   Code/Utils/SI.py:38-67,147-192 ----------------
 * alpha_i = |sum_j b_ij (u_j - u_i)| / max(sum_j |b_ij||u_j - u_i|, floor) over the unit stiffness
 * matrix b, psi = 1/(1+exp(-20(alpha-0.5))), eps_i = psi Cm h_i ||f'(u_i)|| (flux BURGERS/KPP) or
 * psi Cm h_i ||w_i|| (flux ADVECTION).  use_bc != 0: b carries identity Dirichlet rows/cols, as the
 * reference assembles it (Exact_Burger_SI.py:169-172).  psi_out may be NULL. */
int cfem_si_epsilon(cfem_ctx* ctx, int flux, double Cm, double floor, int use_bc, const double* u_n,
                    const double* h, const double* w, double* psi_out, double* eps_out);

/* ---- (a-7, a-8) assembly ----------------------------------------------
 * Matrices are written into the context (CFEM_MAT_SYSTEM); fetch values with
 * cfem_matrix_values.  Vectors go to caller memory.
 * bc_values: one value per Dirichlet dof in the order of the current
 * Dirichlet set (default: sorted boundary dofs), or NULL for 0. */
/* A = M + dt/2 C_w + dt/2 K_eps (bc rows/cols -> identity),
 * b = (M - dt/2 C_w - dt/2 K_eps) u_n, lifted, b_bc = g.   RV_node.py:220-242 */
int cfem_assemble_advection(cfem_ctx* ctx, double dt, const double* w, const double* eps /*NULL: GFEM*/,
                            const double* u_n, const double* bc_values, double* b_out);
/* F(uh) of KPP_exact.py:141-145 / Exact_Burger_RV.py:207-211 with dolfinx
 * NonlinearProblem.F bc handling (lifting x0=uh, alpha=-1; F_bc = uh - g). */
int cfem_assemble_cn_residual(cfem_ctx* ctx, int flux, double dt, const double* uh,
                              const double* u_n, const double* eps, const double* bc_values,
                              double* F_out);
/* J = dF/duh with bc rows/cols -> identity -> CFEM_MAT_SYSTEM. */
int cfem_assemble_cn_jacobian(cfem_ctx* ctx, int flux, double dt, const double* uh,
                              const double* eps);
/* K_eps = int eps grad u . grad v (eps NULL: 1) -> CFEM_MAT_STIFFNESS, no bc. */
int cfem_assemble_stiffness(cfem_ctx* ctx, const double* eps);
/* values of a context matrix in the layout of cfem_get_csr_pattern (nnz doubles) */
int cfem_matrix_values(cfem_ctx* ctx, int which, double* vals_out);

/* ---- (a-9) sparse kernels ---------------------------------------------
 * y = A x  (fp64 CSR, sub-warp per row). */
int cfem_spmv(cfem_ctx* ctx, int which, const double* x, double* y);
/* Jacobi-preconditioned Krylov solve A x = b replacing KSP preonly + PC lu
 * (RV_node.py:131-134, helpers.py:35, NewtonSolver default).  x_io: initial
 * guess in, solution out.  Stops at ||r||_2 <= rtol*||b||_2 (or atol). */
int cfem_solve(cfem_ctx* ctx, int which, int solver, const double* b, double* x_io,
               double rtol, double atol, int max_it, int* iters, double* relres);

/* ---- (a-10) time loops -------------------------------------------------*/
typedef struct cfem_step_params {
  int32_t flux;          /* CFEM_FLUX_*                                          */
  int32_t scheme;        /* residual time derivative: CFEM_BDF1 / CFEM_BDF2      */
  double dt;
  double Cvel, Crv;      /* RV(Cvel, Crv, domain)                   RV.py:7      */
  double newton_rtol;    /* 1e-4   KPP_exact.py:150                              */
  double newton_atol;    /* 1e-10  dolfinx NewtonSolver default                  */
  int32_t newton_max_it; /* 100    KPP_exact.py:149                              */
  int32_t solver;        /* CFEM_SOLVER_* for the non-symmetric systems          */
  double lin_rtol;       /* Krylov ||D^-1 r||/||D^-1 b||, stands in for LU (default 1e-13) */
  int32_t lin_max_it;
  int32_t bc_kind;       /* CFEM_BC_*                                            */
  double bc_value;       /* CFEM_BC_CONSTANT                                     */
  int32_t residual_bc;   /* advection: project the residual with bc (RV_node.py:213) or without (RV_node_convergence.py:188) */
  int32_t mass_solver;   /* CFEM_SOLVER_CHEBYSHEV (default, 0 maps to it) or CFEM_SOLVER_PCG+100 */
  double mass_rtol;      /* tolerance of the residual-projection (mass) solves; 0 = lin_rtol.  All tolerances are on the
                            row-equilibrated residual ||D^-1 r|| / ||D^-1 b|| (D = diag), see DESIGN.md section 4 */
} cfem_step_params;

typedef struct cfem_step_stats {
  int64_t steps;
  int64_t newton_iterations;     /* total over the call                 */
  int64_t mass_iterations;       /* PCG iterations in residual solves   */
  int64_t krylov_iterations;     /* iterations in CN / Jacobian solves  */
  int64_t spmv_launches;
  int64_t assembly_launches;
  int64_t kernel_launches;       /* every kernel of ours in the call    */
  double  last_newton_residual;
  double  time;                  /* simulation time after the call      */
  double  device_ms;             /* CUDA-event time of the whole call on the context stream */
} cfem_step_stats;

/* Load / read the time-loop state (caller numbering, host or device).
 * Any pointer may be NULL (left unchanged / not read).  h = nodal mesh size
 * (cfem_nodal_h output, or caller supplied); w = advection velocity (Nn,2). */
int cfem_state_set(cfem_ctx* ctx, const double* uh, const double* u_n, const double* u_old,
                   const double* u_oo, const double* RH, const double* h, const double* w, double t);
/* same as cfem_state_set but keeps the solvers' iteration-count predictions (a caller that re-sends its
 * host-resident fields every step); cfem_state_set restarts them so that a run is a pure function of its inputs */
int cfem_state_update(cfem_ctx* ctx, const double* uh, const double* u_n, const double* u_old,
                      const double* u_oo, const double* RH, const double* h, const double* w, double t);
int cfem_state_get(cfem_ctx* ctx, double* uh, double* u_n, double* u_old, double* u_oo,
                   double* RH, double* eps, double* t);
/* Distributed hot loops without global arrays: the same two calls with the fields in THIS RANK'S owned
 * order -- entry i belongs to caller dof ordering[i], i < cfem_num_owned (cfem_get_ordering), the layout a
 * rank of an MPI run keeps its fields in.  One contiguous copy per field (host or device pointers), ghosts
 * refreshed by a halo exchange; no permutation, no global-sized buffer.  Keeps the iteration predictions. */
int cfem_state_update_owned(cfem_ctx* ctx, const double* uh, const double* u_n, const double* u_old,
                            const double* u_oo, const double* RH, double t);
int cfem_state_get_owned(cfem_ctx* ctx, double* uh, double* u_n, double* u_old, double* u_oo,
                         double* RH, double* eps, double* t);

/* n_steps passes of the loop body of KPP_exact.py:119-161 /
 * Exact_Burger_RV.py:170-224 on the resident state.  bc_values (CFEM_BC_USER):
 * n_steps * num_dirichlet values.  Returns -3 if Newton fails (dolfinx raises). */
int cfem_step_scalar(cfem_ctx* ctx, const cfem_step_params* p, int n_steps,
                     const double* bc_values, cfem_step_stats* stats);
/* Linear advection: first_gfem != 0 does the plain CN step of
 * RV_node.py:140-157 first; then RV steps of RV_node.py:206-251. */
/* ---- (f-1, f-4) smoothness-indicator stepper + smooth_vector post-filter --------------------------------
 * cfem_step_scalar_si: the loop body of Code/Burgers_equation/Exact_Burger_SI.py:159-197 -- per step the SI
 * viscosity eps = psi(alpha) Cm h ||f'(u_n)|| from the bc'd unit stiffness matrix (SI.py:38-67), the same
 * Crank-Nicolson Newton solve as cfem_step_scalar, then (smooth_l > 0) helpers.smooth_vector(uh, patches,
 * smooth_l) before the rotation.  Cvel / Crv / scheme of the params are unused.
 * cfem_smooth_vector: helpers.smooth_vector (helpers.py:40-50) on its own, in place on u_io (caller numbering):
 * an in-place sweep, node after node in `order` (caller dof ids in sweep order -- the key order of the
 * reference's patches dict; NULL = ascending), each node seeing the already smoothed values of the neighbours
 * that precede it.  Run on the device level by level (exactly the sequential result up to the order of the
 * additions inside one patch sum).  Single-GPU contexts only. */
int cfem_step_scalar_si(cfem_ctx* ctx, const cfem_step_params* p, double Cm, double floor, double smooth_l,
                        const int32_t* smooth_order, int n_steps, const double* bc_values, cfem_step_stats* stats);
int cfem_smooth_vector(cfem_ctx* ctx, double* u_io, const int32_t* order, double l);

int cfem_step_advection(cfem_ctx* ctx, const cfem_step_params* p, int n_steps, int first_gfem,
                        cfem_step_stats* stats);

/* ---- (a-12) compressible Euler, 4-component P1 system -------------------------------------
 * The reference's Code/Compressible_euler/euler_RV.py is a skeleton without an RV term; the
 * scheme is defined by this repository (oracle/euler.py): conserved state
 * U = (rho, m1, m2, E), gamma = 1.4 (euler_RV.py:33,66-72), group-FEM fluxes, BDF2 residual
 * projection, eps = min(Cvel h max_P(|u|+c), Crv h^2 max_k max_P|R_k| / n_k), Crank-Nicolson +
 * Newton (params: dt, Cvel, Crv, newton_*, lin_*), all components Dirichlet on the boundary.
 * Arrays are (Nn,4) row-major in caller numbering; bc_state holds the Dirichlet values (read at
 * the Dirichlet dofs); outputs must be host arrays. */
int cfem_euler_state_set(cfem_ctx* ctx, const double* Uh, const double* Un, const double* Uold, const double* Uoo,
                         const double* bc_state, const double* h, double t);
int cfem_euler_state_get(cfem_ctx* ctx, double* Uh, double* R, double* eps, double* t);
int cfem_step_euler(cfem_ctx* ctx, const cfem_step_params* p, int n_steps, cfem_step_stats* stats);

/* ---- measurement hooks --------------------------------------------------
 * Average device time (ms, CUDA events on the context stream) of `reps`
 * back-to-back launches of one hot kernel on the resident state, and the
 * algorithmic bytes one launch moves. */
enum { CFEM_KERNEL_SPMV = 0, CFEM_KERNEL_ASM_RESIDUAL = 1, CFEM_KERNEL_ASM_JACOBIAN = 2,
       CFEM_KERNEL_RV_EPSILON = 3, CFEM_KERNEL_ASM_RV_RHS = 4, CFEM_KERNEL_PCG_ITER = 5,
       CFEM_KERNEL_COMM_ALLREDUCE = 6 /* 3-scalar all-reduce over the ranks */, CFEM_KERNEL_COMM_HALO = 7 /* full halo exchange of one field */,
       CFEM_KERNEL_SPMV_SYSTEM = 8 /* SpMV with two fused dots on CFEM_MAT_SYSTEM (the BiCGStab kernel) */,
       CFEM_KERNEL_CHEB_ITER = 9 /* a 24-iteration Chebyshev mass solve on CFEM_MAT_MASS_BC (one "launch" = the solve) */,
       CFEM_KERNEL_KRYLOV_ITER = 10 /* a 16-iteration BiCGStab solve on CFEM_MAT_SYSTEM (one "launch" = the solve) */ };
/* Bracket every kernel launch of the following calls with CUDA events (adds ~2 us
 * per launch; use on a separate pass, not on the timed one).  cfem_profile_end sums
 * the device time and launch count per category:
 *   0 SpMV  1 vector assembly  2 matrix assembly  3 Krylov vector kernels
 *   4 RV (stats + epsilon)     5 misc (gather/fill/axpy/bc)
 *   6 fused Chebyshev mass-solve iteration (SpMV + update)
 *   7 communication (halo pack + NCCL send/recv, all-reduce)
 *   8 persistent BiCGStab kernel (one launch = one whole linear solve)  (arrays of 9) */
int cfem_profile_begin(cfem_ctx* ctx, int max_launches);
int cfem_profile_end(cfem_ctx* ctx, double* ms_per_category, int64_t* launches_per_category);
/* Stream idle time between consecutive profiled scopes, attributed to the category of the earlier one (array of 9);
 * call before cfem_profile_end.  Host syncs and launch latency show up here, not in the per-category times. */
int cfem_profile_gaps(cfem_ctx* ctx, double* gap_ms_after_category);
/* L2 error against a P3 interpolant of the exact solution (f-2):  sqrt( int (uh - I3 u_ex)^2 dx ), the functional of
 * Code/Burgers_equation/Exact_Burger_RV_conv.py:81-86,223 and Code/Linear_advection/RV_node_convergence.py:49,69-70,239
 * (u_exact = Function(P3).interpolate(exact); assemble_scalar((uh - u_exact)**2 * dx)).  uh: P1 nodal values (NULL: the
 * resident uh); uex_cells: (n_cells, 10) values of the exact solution at the P3 Lagrange nodes of every CALLER cell, in
 * the order 3 vertices (the cell's own vertex order), 2 nodes on the edge opposite vertex 0, 1, 2 (each one third and
 * two thirds of the way, see cfem_b200/context.py: p3_cell_points), centroid.  Distributed: every rank passes the whole
 * table and gets the global value.  Evaluated in closed form with the 10 x 10 P3 mass matrix (the integrand has
 * degree 6; dolfinx uses a 12-point degree-6 rule -- same value). */
int cfem_l2_error_p3(cfem_ctx* ctx, const double* uh, const double* uex_cells, double* err_out);
int cfem_time_kernel(cfem_ctx* ctx, int kernel, int flux, int reps, double* ms_per_launch,
                     double* algorithmic_bytes);

/* ---- host-side mesh analysis, inspectable without a GPU ------------------
 * The once-per-mesh analysis cfem_create runs (ordering, adjacency, CSR pattern,
 * boundary, tiles, packed codes), exposed so it can be verified on a CPU-only
 * machine.  No arithmetic of the hot path happens here. */
typedef struct cfem_host_mesh cfem_host_mesh;
enum { CFEM_HM_N2U = 0, CFEM_HM_CELLS = 1, CFEM_HM_ROWPTR = 2, CFEM_HM_COLIDX = 3, CFEM_HM_V2C_PTR = 4,
       CFEM_HM_V2C_CODE = 5, CFEM_HM_TILE_NODE = 6, CFEM_HM_TILE_CELLPTR = 7, CFEM_HM_TILE_CELLS = 8,
       CFEM_HM_IS_BND = 9 /* uint8 */, CFEM_HM_BND_USER = 10,
       /* partition (cfem_host_analyse_part): peers and halo lists of this rank */
       CFEM_HM_PEER_RANK = 11, CFEM_HM_SEND_PTR = 12, CFEM_HM_SEND_IDX = 13, CFEM_HM_RECV_OFF = 14,
       CFEM_HM_RECV_CNT = 15,
       CFEM_HM_LAST_CELL = 16 /* per owned node: incident local cell with the highest caller index */,
       /* T16 tile format of the SpMV-type kernels: per CSR entry a 16-bit tile-local column (< 256: row n0 + index of
        * the same tile; otherwise 256 + position in the tile's ascending list of external columns) */
       CFEM_HM_LC16 = 17 /* uint16 */, CFEM_HM_TILE_EXTPTR = 18, CFEM_HM_TILE_EXT = 19,
       CFEM_HM_TILE_ORDER = 20 /* tiles without ghost columns first */ };
int cfem_host_analyse(cfem_host_mesh** out, int64_t n_nodes, int64_t n_cells, const double* x, int xdim,
                      const void* cells, int cell_index_bytes, int order);
/* same, restricted to rank's part of a world-way partition (what cfem_create_distributed builds) */
int cfem_host_analyse_part(cfem_host_mesh** out, int rank, int world, int64_t n_nodes, int64_t n_cells,
                           const double* x, int xdim, const void* cells, int cell_index_bytes, int order);
int cfem_host_analyse_partitioned(cfem_host_mesh** out, int rank, int world, int64_t n_nodes, int64_t n_cells,
                                  const double* x, int xdim, const void* cells, int cell_index_bytes, int order,
                                  const int32_t* node_part);
/* what: 0 owned nodes, 1 local nodes (owned + ghosts), 2 global nodes, 3 local cells, 4 nnz of owned rows */
int64_t cfem_host_info(const cfem_host_mesh* hm, int what);
/* number of elements of array `what` (4-byte elements except CFEM_HM_IS_BND: 1 byte, CFEM_HM_LC16: 2 bytes) */
int64_t cfem_host_size(const cfem_host_mesh* hm, int what);
int cfem_host_copy(const cfem_host_mesh* hm, int what, void* dst);
void cfem_host_free(cfem_host_mesh* hm);

#ifdef __cplusplus
}
#endif
#endif /* CFEM_B200_H */
