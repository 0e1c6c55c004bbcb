// Host-callable launchers implemented in assembly.cu / linalg.cu / rv.cu.
// All pointers are device pointers in INTERNAL (Hilbert) numbering.
#pragma once
#include <functional>

#include "internal.h"

namespace cfem {

// ---- assembly (assembly.cu) --------------------------------------------------
void launch_mass(cfem_ctx* c, Matrix& M, bool bc);
void launch_stiffness(cfem_ctx* c, Matrix& K, const double* eps);
void launch_nodal_h_rhs(cfem_ctx* c, double* b);
// b = M D_t u + flux(u_n) (zeroed on bc dofs if use_bc); fluxn = flux(u_n) (may be null)
void launch_rv_rhs(cfem_ctx* c, int flux, int scheme, double dt, const double* u_n,
                   const double* u_old, const double* u_oo, const double2* w, bool use_bc,
                   double* b, double* fluxn);
// F(uh); fluxn (nullable): precomputed nodal flux(u_n) (else recomputed per cell).
// norm2_partials: per-CTA partial sums of F_i^2 (count returned).
int launch_cn_residual(cfem_ctx* c, int flux, double dt, const double* uh, const double* u_n,
                       const double* eps, const double* g, const double* fluxn, double* F,
                       double* norm2_partials);
void launch_cn_jacobian(cfem_ctx* c, int flux, double dt, const double* uh, const double* eps,
                        Matrix& J);
// F(uh) and J(uh) in one cell pass (first Newton iteration of a step)
int launch_cn_residual_jacobian(cfem_ctx* c, int flux, double dt, const double* uh, const double* u_n,
                                const double* eps, const double* g, const double* fluxn, double* F,
                                double* norm2_partials, Matrix& J);
// A = M + S, b = (M - S) u_n with lifting, S = dt/2 (C_w + K_eps); eps nullable.
void launch_adv_system(cfem_ctx* c, double dt, const double2* w, const double* eps,
                       const double* u_n, const double* g, Matrix& A, double* b);
int assembly_grid(const cfem_ctx* c);
void launch_grad_matrix(cfem_ctx* c, int d, Matrix& C);                                  // int phi_a d_d phi_b
void launch_mass_stiff(cfem_ctx* c, const double* eps, double coef, Matrix& S);          // M + coef K_eps

// ---- smooth_vector post-filter (smooth.cu) -------------------------------------
// in-place sweep over u (internal numbering) in the caller's order (host array of caller dof ids, or null = ascending)
void launch_smooth_vector(cfem_ctx* c, double* u, const int32_t* order_host, double l);
void smooth_plan_free(cfem_ctx* c);

// ---- linear algebra (linalg.cu) ------------------------------------------------
void launch_spmv(cfem_ctx* c, const Matrix& A, const double* x, double* y);
// y = A x with the two fused dot products (y, d0) and (y, y) of a BiCGStab iteration (measurement hook)
void launch_spmv_dots2(cfem_ctx* c, const Matrix& A, const double* x, double* y, const double* d0, double* p0, double* p1);
void launch_gather(cfem_ctx* c, const double* src, const int32_t* idx, double* dst, int64_t n);      // dst[i] = src[idx[i]]
void launch_scatter(cfem_ctx* c, const double* src, const int32_t* idx, double* dst, int64_t n);     // dst[idx[i]] = src[i]
void launch_gather2(cfem_ctx* c, const double2* src, const int32_t* idx, double2* dst, int64_t n);
void launch_gather4(cfem_ctx* c, const double4* src, const int32_t* idx, double4* dst, int64_t n);
void launch_fill(cfem_ctx* c, double* dst, double v, int64_t n);
void launch_copy(cfem_ctx* c, double* dst, const double* src, int64_t n);
// dst[idx[k]] = src[idx[k]], k < n
void launch_copy_indexed(cfem_ctx* c, double* dst, const double* src, const int32_t* idx, int64_t n);
// x -= dx
void launch_sub(cfem_ctx* c, double* x, const double* dx, int64_t n);
struct SolveResult { int iters; double relres; bool converged; };
SolveResult pcg(cfem_ctx* c, const Matrix& A, const double* b, double* x, double rtol, double atol,
                int max_it, int* predict);
// Chebyshev semi-iteration for the (Dirichlet-reduced) P1 mass matrix, spectrum of D^-1 M in [1/2, 2]
SolveResult chebyshev_mass(cfem_ctx* c, const Matrix& A, const double* b, double* x, double rtol, int max_it,
                           int* predict);
SolveResult bicgstab(cfem_ctx* c, const Matrix& A, const double* b, double* x, double rtol,
                     double atol, int max_it, int* predict);
// operator-generic BiCGStab (right Jacobi): apply(x, y, ndot, d0, d1, part0, part1, gated) computes y = A x
// (refreshing the ghosts of x first), writes ndot fused dot-product partials and returns their count.
using LinApply = std::function<int(const double*, double*, int, const double*, const double*, double*, double*, bool)>;
SolveResult bicgstab_generic(cfem_ctx* c, int64_t n, int halo_width, const double* dinv, const LinApply& apply,
                             double* const* work /*8 vectors*/, const double* b, double* x, double rtol, double atol,
                             int max_it, int* predict);
SolveResult gmres(cfem_ctx* c, const Matrix& A, const double* b, double* x, double rtol,
                  double atol, int max_it, int* predict);
double norm2(cfem_ctx* c, const double* v, int64_t n);  // synchronous
// Attach the access-policy window of matrix A (values + pattern -> persisting L2 lines) to the context stream;
// no-op for matrices outside the hot block or when CFEM_L2PERSIST=0.
void l2_prefer(cfem_ctx* c, const Matrix& A);
// detach the window and give the persisting set-aside back to the device
void l2_release(cfem_ctx* c);

// ---- multi-GPU (comm.cu); every call is a no-op when world == 1 --------------------
void comm_unique_id(void* out128);
void comm_init(cfem_ctx* c, int rank, int world, const void* id128);
void comm_destroy(cfem_ctx* c);
void comm_setup_exchange(cfem_ctx* c);   // after the device arrays exist: maps the peer mailboxes (CUDA IPC)
void comm_check(cfem_ctx* c);            // throws if a peer-memory exchange timed out
void halo_exchange(cfem_ctx* c, double* v, int width = 1);   // width doubles per node (1, 2 or 4)
struct GhostSrc;
// producer half only (see p2p.cuh); in_consumer: do not launch anything, the consumer's CTA 0 pushes
GhostSrc halo_push(cfem_ctx* c, double* v, bool gated, bool in_consumer = false);
// local reduce of each partial array to its element 0 + all-reduce; returns the partial count to use after
int allreduce_partials(cfem_ctx* c, int nslots, double* const* slots, const int* ops /*0 sum,1 min,2 max*/, int npart);
int allreduce_sum1(cfem_ctx* c, double* slot, int npart);
// in-kernel finalisation of reductions (p2p.cuh): descriptor for the NEXT producer launch -- takes a fresh all-reduce
// sequence number in a distributed context; fin_available: one GPU, or the peer-memory path is up
struct Fin;
Fin make_fin(cfem_ctx* c);
bool fin_available(const cfem_ctx* c);

// ---- persistent BiCGStab (persist.cu): one cooperative launch for all iterations of a solve
bool bicgstab_persist_available(cfem_ctx* c);
void launch_bicg_persist(cfem_ctx* c, const Matrix& A, const double* rhat, double* x, double* r, double* p, double* v,
                         double* t, double rtol2, double atol2, int max_it);
void persist_plan_free(cfem_ctx* c);
bool bicgstab_async_available(cfem_ctx* c);
void bicgstab_persist_begin(cfem_ctx* c, const Matrix& A, const double* b, double* x, double rtol, double atol, int max_it);
SolveResult bicgstab_persist_end(cfem_ctx* c);   // after a stream synchronisation that follows _begin
void persist_seq_reserve(cfem_ctx* c, int64_t halo_exchanges, int64_t allreduces);   // protocol sequence numbers only
void persist_comm_count(cfem_ctx* c, int64_t halo_exchanges, int64_t allreduces);    // statistics only
void launch_sub_unless_below(cfem_ctx* c, double* x, const double* dx, int64_t n, const double* norm2, double thresh2);
bool cheb_persist_available(cfem_ctx* c);
void launch_cheb_persist(cfem_ctx* c, const Matrix& A, const double* b, double* x_in, double* x_other, double* d,
                         bool first, int iters, double rho0, double sigma1, double theta, double delta);
struct P2PDev;
// exchange tables / current sequence numbers for a kernel that runs several exchanges by itself, and the bump of the
// host-side sequence counters once the number of iterations it ran is known (comm.cu; no-ops on one GPU)
void persist_comm_args(cfem_ctx* c, const P2PDev** dev, const char** mailbox, size_t* halo_off, size_t* halo_stride,
                       const int32_t** peer_rank, int* npeer, int** error, unsigned long long* halo_seq,
                       unsigned long long* red_seq, unsigned long long** tim);
void persist_comm_advance(cfem_ctx* c, int64_t halo_exchanges, int64_t allreduces);
void comm_timers(cfem_ctx* c, double* out8, bool reset);

// ---- RV (rv.cu) ------------------------------------------------------------------
// sum / min / max of v -> c->scalars[0..2] (device), no host sync
void launch_stats(cfem_ctx* c, const double* v);
void launch_epsilon(cfem_ctx* c, int variant, int flux, double Cvel, double Crv, const double* uh,
                    const double* u_n, double* Rh, const double* h, const double2* w, double* eps);
void launch_si_epsilon(cfem_ctx* c, int flux, double Cm, double floor_, bool use_bc, const Matrix& K1, const double* u,
                       const double* h, const double2* w, double* psi, double* eps);
// sum over this rank's cells (each cell counted by one rank) and over the ranks of  int_K (uh - I3 u_ex)^2 ; synchronous.
// table: the exact solution at the ten P3 nodes per cell, rows by local cell (local_rows) or by caller cell index
double launch_l2_error_p3(cfem_ctx* c, const double* uh, const double* table, bool local_rows);
void launch_bc_values(cfem_ctx* c, int kind, double value, double t, const double* user_vals,
                      double* g);

// ---- Euler system (euler.cu) -------------------------------------------------------------
void euler_state_ptrs(cfem_ctx* c, double** Uh, double** Un, double** Uold, double** Uoo, double** G, double** R);
void euler_reset_predictions(cfem_ctx* c);
void euler_free(cfem_ctx* c);
void euler_steps(cfem_ctx* c, const cfem_step_params* p, int n_steps, cfem_step_stats* st);

}  // namespace cfem
