// Fused residual-viscosity kernels: global sum/min/max of a nodal field (one
// pass, fixed-order two-stage reduction), the patch max/min + epsilon kernel
// (sub-warp per node, shuffle reductions over the node's CSR row == its patch),
// and the Dirichlet-value kernel.
//
// Restates the per-node Python loops of the reference, Code/Utils/RV.py:27-142.
#include "device_utils.cuh"
#include "launch.h"

namespace cfem {

#define LAUNCHED(c) do { CUDA_OK(cudaGetLastError()); (c)->launches.total++; } while (0)

static inline int vec_grid(const cfem_ctx* c, int64_t n) {
  int64_t b = (n + kBlock - 1) / kBlock;
  const int64_t cap = (int64_t)c->sm_count * 8;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

enum { P_SUM = 5, P_MIN = 6, P_MAX = 7 };  // partial slots (shared with linalg.cu's P_A..P_C)

// ||f'(u)||_2 with the operation order of np.linalg.norm(np.array(f'(u)))
template <int FLUX>
__device__ __forceinline__ double beta_of(double u) {
  if (FLUX == CFEM_FLUX_BURGERS) return sqrt(__dadd_rn(__dmul_rn(u, u), __dmul_rn(u, u)));
  double s, c;
  sincos(u, &s, &c);
  return sqrt(__dadd_rn(__dmul_rn(c, c), __dmul_rn(s, s)));
}

// one pass: partial sum / min / max of v; optionally beta[i] = ||f'(v_i)||
template <int FLUX>
__global__ void __launch_bounds__(kBlock)
k_stats(int64_t n_owned, int64_t n_local, const double* __restrict__ v, double* __restrict__ beta,
        double* __restrict__ part) {
  pdl_wait();
  pdl_launch();
  __shared__ double red[9];
  double s = 0.0, mn = INFINITY, mx = -INFINITY;
  const int64_t n = FLUX >= 0 ? n_local : n_owned;  // beta is also needed on the ghosts
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const double x = v[i];
    if (i < n_owned) { s += x; mn = fmin(mn, x); mx = fmax(mx, x); }
    if (FLUX >= 0) beta[i] = beta_of<FLUX>(x);
  }
  s = block_sum(s, red); mn = block_min(mn, red); mx = block_max(mx, red);
  if (threadIdx.x == 0) {
    part[P_SUM * kMaxPartials + blockIdx.x] = s;
    part[P_MIN * kMaxPartials + blockIdx.x] = mn;
    part[P_MAX * kMaxPartials + blockIdx.x] = mx;
  }
}

// ||v - mean(v)||_inf from the partials (np.linalg.norm(u - np.mean(u), ord=inf), RV.py:59)
__device__ __forceinline__ double absolute_term(const double* part, int npart, int64_t n, double* red) {
  double s = 0.0, mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < npart; i += kBlock) {
    s += part[P_SUM * kMaxPartials + i];
    mn = fmin(mn, part[P_MIN * kMaxPartials + i]);
    mx = fmax(mx, part[P_MAX * kMaxPartials + i]);
  }
  s = block_sum(s, red); mn = block_min(mn, red); mx = block_max(mx, red);
  const double mean = s / (double)n;
  return fmax(fabs(mx - mean), fabs(mn - mean));
}

// Python min(a, b): b if b < a else a  (NaN / inf second argument keeps a)
__device__ __forceinline__ double pymin(double a, double b) { return b < a ? b : a; }

// RV.get_epsilon_nonlinear (RV.py:56-90) / get_epsilon_linear (RV.py:92-127).
// Tile-streamed: the tile's column indices are staged in
// shared memory once (coalesced), then three coalesced gather passes (u_n, |Rh|, beta)
// park values in shared memory and thread r reduces row r's patch.
template <bool LINEAR>
__global__ void __launch_bounds__(kTileNodes)
k_epsilon_stream(const int ntiles, const int64_t nn, const int32_t* __restrict__ tile_node,
                 const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                 const double* __restrict__ u_n, const double* __restrict__ Rh, const double* __restrict__ beta,
                 const double2* __restrict__ w, const double* __restrict__ h, const double* __restrict__ part,
                 int npart, double Cvel, double Crv, double* __restrict__ eps) {
  pdl_wait();
  pdl_launch();
  __shared__ double val[kTileNnzCap];
  __shared__ int32_t col[kTileNnzCap];
  __shared__ int32_t rp[kTileNodes + 1];
  __shared__ double red[9];
  const double A = absolute_term(part, npart, nn, red);
  const int tid = threadIdx.x;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n0 = tile_node[tile], nrows = tile_node[tile + 1] - n0;
    for (int i = tid; i <= nrows; i += kTileNodes) rp[i] = rowptr[n0 + i];
    __syncthreads();
    const int start = rp[0], cnt = rp[nrows] - start;
    for (int p = tid; p < cnt; p += kTileNodes) { const int j = colidx[start + p]; col[p] = j; val[p] = u_n[j]; }
    __syncthreads();
    const int a = tid < nrows ? rp[tid] - start : 0, e = tid < nrows ? rp[tid + 1] - start : 0;
    double umax = -INFINITY, umin = INFINITY, rmax = 0.0, bmax = 0.0;
    for (int k = a; k < e; ++k) { umax = fmax(umax, val[k]); umin = fmin(umin, val[k]); }
    __syncthreads();
    for (int p = tid; p < cnt; p += kTileNodes) val[p] = fabs(Rh[col[p]]);
    __syncthreads();
    for (int k = a; k < e; ++k) rmax = fmax(rmax, val[k]);
    if (!LINEAR) {
      __syncthreads();
      for (int p = tid; p < cnt; p += kTileNodes) val[p] = beta[col[p]];
      __syncthreads();
      for (int k = a; k < e; ++k) bmax = fmax(bmax, val[k]);
    }
    if (tid < nrows) {
      const int row = n0 + tid;
      if (LINEAR) {
        const double2 wi = w[row];  // centre node, RV.py:113-115
        bmax = sqrt(__dadd_rn(__dmul_rn(wi.x, wi.x), __dmul_rn(wi.y, wi.y)));
      }
      const double hi = h[row];
      const double n_i = fabs((umax - umin) - A);
      const double Ri = rmax / n_i;
      const double first = __dmul_rn(__dmul_rn(Cvel, hi), bmax);
      const double second = __dmul_rn(__dmul_rn(Crv, __dmul_rn(hi, hi)), fabs(Ri));
      eps[row] = pymin(first, second);
    }
    __syncthreads();
  }
}

// pointwise variants: RV.get_epsilon (RV.py:27-40), get_epsilon_1storder (RV.py:42-54),
// get_epsilon_linear_simple (RV.py:129-142; also normalises Rh in place)
template <int FLUX>
__global__ void __launch_bounds__(kBlock)
k_epsilon_pointwise(int64_t n, int64_t n_global, int variant, const double* __restrict__ uh, double* __restrict__ Rh,
                    const double* __restrict__ h, const double2* __restrict__ w, const double* __restrict__ part,
                    int npart, double Cvel, double Crv, double* __restrict__ eps) {
  __shared__ double red[9];
  double A = 1.0;
  if (variant == CFEM_EPS_LINEAR_SIMPLE) A = absolute_term(part, npart, n_global, red);
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    double b;
    if (FLUX == CFEM_FLUX_ADVECTION) {
      const double2 wi = w[i];
      b = sqrt(__dadd_rn(__dmul_rn(wi.x, wi.x), __dmul_rn(wi.y, wi.y)));
    } else {
      b = beta_of<FLUX == CFEM_FLUX_ADVECTION ? CFEM_FLUX_BURGERS : FLUX>(uh[i]);
    }
    const double hi = h[i];
    if (variant == CFEM_EPS_FIRST_ORDER) {
      eps[i] = __dmul_rn(__dmul_rn(0.5, hi), b);
    } else {
      double r = Rh[i];
      if (variant == CFEM_EPS_LINEAR_SIMPLE) { r = r / A; Rh[i] = r; }
      eps[i] = pymin(__dmul_rn(__dmul_rn(Cvel, hi), b), __dmul_rn(__dmul_rn(Crv, __dmul_rn(hi, hi)), fabs(r)));
    }
  }
}

// Cell-based viscosity of Code/Linear_advection/RV_cell.py:174-192: the residual is divided by
// max(u_n - mean(u_n)), every cell gets min(Cvel h_K max|w|, Crv h_K^2 max|R|) over its three dofs with
// h_K its shortest edge, and the per-cell loop writes that value to the three dofs in cell order -- so a
// node ends up with the value of its highest-numbered cell (last_cell).  One thread per owned node.
__global__ void __launch_bounds__(kBlock)
k_epsilon_cell(int64_t n_owned, int64_t n_global, const int32_t* __restrict__ cells,
               const int32_t* __restrict__ last_cell, const double2* __restrict__ xy,
               const double* __restrict__ Rh, const double2* __restrict__ w, const double* __restrict__ part,
               int npart, double Cvel, double Crv, double* __restrict__ eps) {
  __shared__ double red[9];
  double s = 0.0, mx = -INFINITY;
  for (int i = threadIdx.x; i < npart; i += kBlock) {
    s += part[P_SUM * kMaxPartials + i];
    mx = fmax(mx, part[P_MAX * kMaxPartials + i]);
  }
  s = block_sum(s, red); mx = block_max(mx, red);
  const double A = mx - s / (double)n_global;   // np.max(u_n - np.mean(u_n))
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n_owned; i += (int64_t)gridDim.x * kBlock) {
    const int64_t c = last_cell[i];
    double2 p[3];
    double Rk = 0.0, Bk = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int32_t v = cells[3 * c + k];
      p[k] = xy[v];
      const double2 wv = w[v];
      Rk = fmax(Rk, fabs(Rh[v] / A));
      Bk = fmax(Bk, sqrt(__dadd_rn(__dmul_rn(wv.x, wv.x), __dmul_rn(wv.y, wv.y))));
    }
    double hk = INFINITY;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = a + 1; b < 3; ++b) {
        const double dx = p[a].x - p[b].x, dy = p[a].y - p[b].y;
        hk = fmin(hk, sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))));
      }
    eps[i] = pymin(__dmul_rn(__dmul_rn(Cvel, hk), Bk), __dmul_rn(__dmul_rn(Crv, __dmul_rn(hk, hk)), Rk));
  }
}

// all-reduce the sum / min / max partials over the ranks; returns the partial count to use
static int stats_allreduce(cfem_ctx* c, int gv) {
  double* sl[3] = {c->partials + P_SUM * kMaxPartials, c->partials + P_MIN * kMaxPartials, c->partials + P_MAX * kMaxPartials};
  const int op[3] = {0, 1, 2};
  return allreduce_partials(c, 3, sl, op, gv);
}

void launch_stats(cfem_ctx* c, const double* v) {
  const int gv = vec_grid(c, c->dm.nn);
  launch_pdl(k_stats<-1>, gv, kBlock, 0, c->stream, c->dm.no, c->dm.nn, v, nullptr, c->partials); LAUNCHED(c);
  stats_allreduce(c, gv);
}

void launch_epsilon(cfem_ctx* c, int variant, int flux, double Cvel, double Crv, const double* uh,
                    const double* u_n, double* Rh, const double* h, const double2* w, double* eps) {
  ProfScope ps(c, PROF_RV);
  const int64_t n = c->dm.no, nl = c->dm.nn, ng = c->dm.nn_global;
  const int gv = vec_grid(c, nl);
  int np = gv;
  if (variant == CFEM_EPS_CELL) {
    if (!u_n || !Rh || !w) CFEM_THROW(-1, "rv_epsilon(cell): u_n, Rh and w are required");
    launch_pdl(k_stats<-1>, gv, kBlock, 0, c->stream, n, nl, u_n, nullptr, c->partials); LAUNCHED(c);
    np = stats_allreduce(c, gv);
    k_epsilon_cell<<<vec_grid(c, n), kBlock, 0, c->stream>>>(n, ng, c->dm.cells, c->dm.last_cell, c->dm.xy, Rh, w,
                                                            c->partials, np, Cvel, Crv, eps);
    LAUNCHED(c);
    halo_exchange(c, eps);
    return;
  }
  if (!h) CFEM_THROW(-1, "rv_epsilon: nodal mesh size h is required");
  if (variant == CFEM_EPS_NONLINEAR || variant == CFEM_EPS_LINEAR) {
    if (!uh || !u_n || !Rh) CFEM_THROW(-1, "rv_epsilon: uh, u_n and Rh are required");
    double* beta = c->wk[9];
    if (variant == CFEM_EPS_LINEAR) {
      if (!w) CFEM_THROW(-1, "rv_epsilon(linear): velocity field w is required");
      launch_pdl(k_stats<-1>, gv, kBlock, 0, c->stream, n, nl, uh, nullptr, c->partials);
    } else if (flux == CFEM_FLUX_BURGERS) {
      launch_pdl(k_stats<CFEM_FLUX_BURGERS>, gv, kBlock, 0, c->stream, n, nl, uh, beta, c->partials);
    } else if (flux == CFEM_FLUX_KPP) {
      launch_pdl(k_stats<CFEM_FLUX_KPP>, gv, kBlock, 0, c->stream, n, nl, uh, beta, c->partials);
    } else {
      CFEM_THROW(-1, "rv_epsilon(nonlinear): flux must be BURGERS or KPP");
    }
    LAUNCHED(c);
    np = stats_allreduce(c, gv);
    int64_t g = (int64_t)c->sm_count * 4;
    if (g > c->dm.ntiles) g = c->dm.ntiles;
    if (variant == CFEM_EPS_LINEAR)
      launch_pdl(k_epsilon_stream<true>, (int)g, kTileNodes, 0, c->stream, c->dm.ntiles, ng, c->dm.tile_node, c->dm.rowptr, c->dm.colidx,
                                                                 u_n, Rh, nullptr, w, h, c->partials, np, Cvel, Crv, eps);
    else
      launch_pdl(k_epsilon_stream<false>, (int)g, kTileNodes, 0, c->stream, c->dm.ntiles, ng, c->dm.tile_node, c->dm.rowptr, c->dm.colidx,
                                                                  u_n, Rh, beta, w, h, c->partials, np, Cvel, Crv, eps);
    LAUNCHED(c);
    halo_exchange(c, eps);
    return;
  }
  if (variant == CFEM_EPS_POINTWISE || variant == CFEM_EPS_FIRST_ORDER || variant == CFEM_EPS_LINEAR_SIMPLE) {
    if (variant != CFEM_EPS_FIRST_ORDER && !Rh) CFEM_THROW(-1, "rv_epsilon: residual is required");
    if (variant == CFEM_EPS_LINEAR_SIMPLE) {
      if (!u_n) CFEM_THROW(-1, "rv_epsilon(linear_simple): u_n is required");
      flux = CFEM_FLUX_ADVECTION;
      launch_pdl(k_stats<-1>, gv, kBlock, 0, c->stream, n, nl, u_n, nullptr, c->partials); LAUNCHED(c);
      np = stats_allreduce(c, gv);
    }
    if (flux == CFEM_FLUX_ADVECTION) {
      if (!w) CFEM_THROW(-1, "rv_epsilon: velocity field w is required");
      k_epsilon_pointwise<CFEM_FLUX_ADVECTION><<<gv, kBlock, 0, c->stream>>>(nl, ng, variant, uh, Rh, h, w, c->partials, np, Cvel, Crv, eps);
    } else if (flux == CFEM_FLUX_BURGERS) {
      if (!uh) CFEM_THROW(-1, "rv_epsilon: uh is required");
      k_epsilon_pointwise<CFEM_FLUX_BURGERS><<<gv, kBlock, 0, c->stream>>>(nl, ng, variant, uh, Rh, h, w, c->partials, np, Cvel, Crv, eps);
    } else if (flux == CFEM_FLUX_KPP) {
      if (!uh) CFEM_THROW(-1, "rv_epsilon: uh is required");
      k_epsilon_pointwise<CFEM_FLUX_KPP><<<gv, kBlock, 0, c->stream>>>(nl, ng, variant, uh, Rh, h, w, c->partials, np, Cvel, Crv, eps);
    } else {
      CFEM_THROW(-1, "rv_epsilon: unknown flux");
    }
    LAUNCHED(c);
    return;
  }
  CFEM_THROW(-1, "rv_epsilon: unknown variant");
}

// ---------------------------------------------------------------- smoothness indicator (SI.py:38-67,147-192)
template <int FLUX>
__global__ void __launch_bounds__(kBlock)
k_si_epsilon(int64_t no, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
             const double* __restrict__ K, const uint8_t* __restrict__ is_bc, const double* __restrict__ u,
             const double* __restrict__ h, const double2* __restrict__ w, double Cm, double floor_,
             double* __restrict__ psi_out, double* __restrict__ eps) {
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < no; i += (int64_t)gridDim.x * kBlock) {
    const double ui = u[i];
    double num = 0.0, den = 0.0;
    if (!(is_bc && is_bc[i])) {  // an identity row only sees itself: both sums stay 0
      const int p1 = rowptr[i + 1];
      for (int p = rowptr[i]; p < p1; ++p) {
        const int j = colidx[p];
        if (is_bc && is_bc[j]) continue;  // zeroed Dirichlet column
        const double du = u[j] - ui, b = K[p];
        num += b * du;
        den += fabs(b) * fabs(du);
      }
    }
    const double alpha = fabs(num) / fmax(den, floor_);
    const double psi = 1.0 / (1.0 + exp(-20.0 * (alpha - 0.5)));
    double fn;
    if (FLUX == CFEM_FLUX_ADVECTION) { const double2 wi = w[i]; fn = sqrt(__dadd_rn(__dmul_rn(wi.x, wi.x), __dmul_rn(wi.y, wi.y))); }
    else fn = beta_of<FLUX == CFEM_FLUX_ADVECTION ? CFEM_FLUX_BURGERS : FLUX>(ui);
    if (psi_out) psi_out[i] = psi;
    eps[i] = psi * Cm * h[i] * fn;
  }
}

void launch_si_epsilon(cfem_ctx* c, int flux, double Cm, double floor_, bool use_bc, const Matrix& K1, const double* u,
                       const double* h, const double2* w, double* psi, double* eps) {
  ProfScope ps(c, PROF_RV);
  const int64_t no = c->dm.no;
  const int g = vec_grid(c, no);
  const uint8_t* bc = use_bc ? c->dm.is_bc : nullptr;
  if (flux == CFEM_FLUX_ADVECTION) {
    if (!w) CFEM_THROW(-1, "si_epsilon: velocity field w is required");
    k_si_epsilon<CFEM_FLUX_ADVECTION><<<g, kBlock, 0, c->stream>>>(no, c->dm.rowptr, c->dm.colidx, K1.vals, bc, u, h, w, Cm, floor_, psi, eps);
  } else if (flux == CFEM_FLUX_BURGERS) {
    k_si_epsilon<CFEM_FLUX_BURGERS><<<g, kBlock, 0, c->stream>>>(no, c->dm.rowptr, c->dm.colidx, K1.vals, bc, u, h, w, Cm, floor_, psi, eps);
  } else if (flux == CFEM_FLUX_KPP) {
    k_si_epsilon<CFEM_FLUX_KPP><<<g, kBlock, 0, c->stream>>>(no, c->dm.rowptr, c->dm.colidx, K1.vals, bc, u, h, w, Cm, floor_, psi, eps);
  } else {
    CFEM_THROW(-1, "si_epsilon: unknown flux");
  }
  LAUNCHED(c);
  halo_exchange(c, eps);
}

// ---------------------------------------------------------------- Dirichlet data
// Exact solution of the 2-D Burgers Riemann problem, same branch order and
// operation order as the reference (Code/Burgers_equation/Exact_Burger_RV.py:37-66),
// unfused arithmetic so that nodes lying on a front classify identically.
__device__ double burgers_exact(double X, double Y, double t) {
  const double half = 0.5;
  double u = 0.0;
  const double a1 = __dsub_rn(half, __ddiv_rn(__dmul_rn(3.0, t), 5.0));   // 1/2 - 3t/5
  const double y1 = __dadd_rn(half, __ddiv_rn(__dmul_rn(3.0, t), 20.0));  // 1/2 + 3t/20
  const bool m1 = X <= a1;
  if (m1 && Y > y1) u = -0.2;
  if (m1 && Y <= y1) u = 0.5;
  const double a2 = __dsub_rn(half, __ddiv_rn(t, 4.0));                   // 1/2 - t/4
  const bool m2 = (a1 <= X) && (X <= a2);
  const double l2 = __dsub_rn(__dadd_rn(__ddiv_rn(__dmul_rn(-8.0, X), 7.0), 15.0 / 14.0),
                              __ddiv_rn(__dmul_rn(15.0, t), 28.0));
  if (m2 && Y > l2) u = -1.0;
  if (m2 && Y <= l2) u = 0.5;
  const double a3 = __dadd_rn(half, __ddiv_rn(t, 2.0));                   // 1/2 + t/2
  const bool m3 = (a2 <= X) && (X <= a3);
  const double l3 = __dsub_rn(__dadd_rn(__ddiv_rn(X, 6.0), 5.0 / 12.0), __ddiv_rn(__dmul_rn(5.0, t), 24.0));
  if (m3 && Y > l3) u = -1.0;
  if (m3 && Y <= l3) u = 0.5;
  const double a4 = __dadd_rn(half, __ddiv_rn(__dmul_rn(4.0, t), 5.0));   // 1/2 + 4t/5
  const bool m4 = (a3 <= X) && (X <= a4);
  const double q = __dsub_rn(__dadd_rn(X, t), half);
  const double l4 = __dsub_rn(X, __dmul_rn(__ddiv_rn(5.0, __dmul_rn(18.0, t)), __dmul_rn(q, q)));
  if (m4 && Y > l4) u = -1.0;
  if (m4 && Y <= l4) u = __ddiv_rn(__dsub_rn(__dmul_rn(2.0, X), 1.0), __dmul_rn(2.0, t));
  const bool m5 = X >= a4;
  const double y5 = __dsub_rn(half, __ddiv_rn(t, 10.0));
  if (m5 && Y > y5) u = -1.0;
  if (m5 && Y <= y5) u = 0.8;
  return u;
}

__global__ void k_bc_values(int64_t nbc, const int32_t* __restrict__ bc_nodes, const int32_t* __restrict__ bc_pos,
                            int kind, double value, double t, const double* __restrict__ user,
                            const double2* __restrict__ xy, double* __restrict__ g) {
  for (int64_t j = blockIdx.x * (int64_t)kBlock + threadIdx.x; j < nbc; j += (int64_t)gridDim.x * kBlock) {
    const int node = bc_nodes[j];
    double v;
    if (kind == CFEM_BC_CONSTANT) v = value;
    else if (kind == CFEM_BC_USER) v = user[bc_pos[j]];
    else { const double2 p = xy[node]; v = burgers_exact(p.x, p.y, t); }
    g[node] = v;
  }
}

void launch_bc_values(cfem_ctx* c, int kind, double value, double t, const double* user_vals, double* g) {
  if (c->nbc == 0) return;
  if (kind == CFEM_BC_USER && !user_vals) CFEM_THROW(-1, "CFEM_BC_USER needs bc_values");
  ProfScope ps(c, PROF_MISC);
  k_bc_values<<<vec_grid(c, c->nbc), kBlock, 0, c->stream>>>(c->nbc, c->d_bc_nodes, c->d_bc_pos, kind, value, t, user_vals, c->dm.xy, g);
  LAUNCHED(c);
}

}  // namespace cfem
