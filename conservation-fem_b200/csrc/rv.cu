// Fused residual-viscosity kernels: global sum/min/max of a nodal field (one
// pass, fixed-order two-stage reduction), the patch max/min + epsilon kernel
// (sub-warp per node, shuffle reductions over the node's CSR row == its patch),
// and the Dirichlet-value kernel.
//
// Restates the per-node Python loops of the reference, Code/Utils/RV.py:27-142.
#include <algorithm>
#include <cstdlib>
#include <string>

#include "device_utils.cuh"
#include "launch.h"
#include "p2p.cuh"

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

// Same formulas, one gather pass (default).  The tile's patch values are staged ONCE per tile in the T16 layout of the
// SpMV-type kernels -- own rows coalesced, the ~100 external columns gathered -- for all three fields at a time
// (u_n, |Rh|, ||f'(uh)|| computed on the fly: no beta table is written or read), the 16-bit tile-local columns are
// staged coalesced, and thread r reduces row r's patch out of shared memory.  Replaces three serial gather passes
// over 32-bit columns with four barriers each (77 us -> see profiles/ at 1 M nodes).
template <bool LINEAR, int FLUX>
__global__ void __launch_bounds__(kTileNodes)
k_epsilon_t16(const int ntiles, const int64_t nn, const int ext_cap, const int32_t* __restrict__ tile_node,
              const int32_t* __restrict__ rowptr, const uint16_t* __restrict__ lc16, const int32_t* __restrict__ extptr,
              const int32_t* __restrict__ ext, const double* __restrict__ uh, const double* __restrict__ u_n,
              const double* __restrict__ Rh, const double2* __restrict__ w, const double* __restrict__ h,
              const double* __restrict__ part, int npart, double Cvel, double Crv, double* __restrict__ eps) {
  pdl_wait();
  pdl_launch();
  extern __shared__ double eps_smem[];
  const int span = kTileNodes + ext_cap;
  double* const s_un = eps_smem;             // [span]
  double* const s_r = eps_smem + span;       // [span]
  double* const s_b = eps_smem + 2 * span;   // [span] (nonlinear variants)
  uint16_t* const s_lc = (uint16_t*)(eps_smem + 3 * span);   // [kTileNnzCap]
  __shared__ int32_t rp[kTileNodes + 1];
  __shared__ double red[9];
  const double A = absolute_term(part, npart, nn, red);
  const int tid = threadIdx.x;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n0 = tile_node[tile], nrows = tile_node[tile + 1] - n0;
    const int e0 = extptr[tile], ne = extptr[tile + 1] - e0;
    const int start = rowptr[n0], cnt = rowptr[n0 + nrows] - start;
    for (int i = tid; i <= nrows; i += kTileNodes) rp[i] = rowptr[n0 + i] - start;
    for (int p = tid; p < cnt; p += kTileNodes) s_lc[p] = lc16[start + p];
    for (int e = tid; e < nrows + ne; e += kTileNodes) {
      const int j = e < nrows ? n0 + e : ext[e0 + e - nrows];
      const int slot = e < nrows ? e : kTileNodes + (e - nrows);
      s_un[slot] = u_n[j];
      s_r[slot] = fabs(Rh[j]);
      if (!LINEAR) s_b[slot] = beta_of<FLUX>(uh[j]);
    }
    __syncthreads();
    if (tid < nrows) {
      double umax = -INFINITY, umin = INFINITY, rmax = 0.0, bmax = 0.0;
      for (int k = rp[tid]; k < rp[tid + 1]; ++k) {
        const int j = s_lc[k];
        const double un = s_un[j];
        umax = fmax(umax, un);
        umin = fmin(umin, un);
        rmax = fmax(rmax, s_r[j]);
        if (!LINEAR) bmax = fmax(bmax, s_b[j]);
      }
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

template <bool LINEAR, int FLUX>
static void launch_epsilon_t16(cfem_ctx* c, const double* uh, const double* u_n, const double* Rh, const double2* w,
                               const double* h, int np, double Cvel, double Crv, double* eps) {
  const DevMesh& m = c->dm;
  const size_t smem = sizeof(double) * 3 * (size_t)(kTileNodes + m.ext_cap) + sizeof(uint16_t) * (size_t)kTileNnzCap;
  if (smem > kDynSmemCeiling) CFEM_THROW(-2, "epsilon kernel: a tile has too many external columns for shared memory");
  CUDA_OK(cudaFuncSetAttribute(k_epsilon_t16<LINEAR, FLUX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynSmemCeiling));
  int64_t g = (int64_t)c->sm_count * 8;
  if (g > m.ntiles) g = m.ntiles;
  launch_pdl(k_epsilon_t16<LINEAR, FLUX>, (int)g, kTileNodes, smem, c->stream, m.ntiles, m.nn_global, m.ext_cap, m.tile_node,
             m.rowptr, m.lc16, m.tile_extptr, m.tile_ext, uh, u_n, Rh, w, h, c->partials, np, Cvel, Crv, eps);
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
    static const bool one_pass = !(getenv("CFEM_EPS") && std::string(getenv("CFEM_EPS")) == "3pass");
    if (one_pass) {
      if (variant == CFEM_EPS_LINEAR && !w) CFEM_THROW(-1, "rv_epsilon(linear): velocity field w is required");
      if (variant == CFEM_EPS_NONLINEAR && flux != CFEM_FLUX_BURGERS && flux != CFEM_FLUX_KPP)
        CFEM_THROW(-1, "rv_epsilon(nonlinear): flux must be BURGERS or KPP");
      launch_pdl(k_stats<-1>, gv, kBlock, 0, c->stream, n, nl, uh, nullptr, c->partials); LAUNCHED(c);
      np = stats_allreduce(c, gv);
      if (variant == CFEM_EPS_LINEAR) launch_epsilon_t16<true, CFEM_FLUX_BURGERS>(c, uh, u_n, Rh, w, h, np, Cvel, Crv, eps);
      else if (flux == CFEM_FLUX_BURGERS) launch_epsilon_t16<false, CFEM_FLUX_BURGERS>(c, uh, u_n, Rh, w, h, np, Cvel, Crv, eps);
      else launch_epsilon_t16<false, CFEM_FLUX_KPP>(c, uh, u_n, Rh, w, h, np, Cvel, Crv, eps);
      LAUNCHED(c);
      halo_exchange(c, eps);
      return;
    }
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

// ---------------------------------------------------------------- (f-2) L2 error against a P3 interpolant
// P3 mass matrix of the reference triangle, int phi_a phi_b / |K| (x 6720; symmetric).  Node order: vertices 0 1 2, edge
// opposite vertex 0 (1/3 and 2/3 of the way from vertex 1 to 2), edge opposite 1 (from 0 to 2), edge opposite 2
// (from 0 to 1), centroid.  Filled once from the closed form  int l0^a l1^b l2^c = 2|K| a! b! c! / (a+b+c+2)! .
__constant__ double kM3[10][10];

static void fill_m3(double M[10][10]) {
  // basis in barycentric monomials: coefficient arrays over exponents (a,b,c) with a+b+c <= 3
  struct Poly { double c[4][4][4]; };
  auto zero = []() { Poly p; for (auto& x : p.c) for (auto& y : x) for (auto& z : y) z = 0.0; return p; };
  auto mul_lin = [&](const Poly& p, int i, double a, double b) {   // p * (a l_i + b)
    Poly r = zero();
    for (int x = 0; x < 4; ++x) for (int y = 0; y < 4; ++y) for (int z = 0; z < 4; ++z) {
      const double v = p.c[x][y][z];
      if (v == 0.0) continue;
      r.c[x][y][z] += b * v;
      const int nx = x + (i == 0), ny = y + (i == 1), nz = z + (i == 2);
      if (nx < 4 && ny < 4 && nz < 4) r.c[nx][ny][nz] += a * v;
    }
    return r;
  };
  Poly one = zero();
  one.c[0][0][0] = 1.0;
  Poly phi[10];
  for (int i = 0; i < 3; ++i) {   // 1/2 l (3l-1)(3l-2)
    phi[i] = mul_lin(mul_lin(mul_lin(one, i, 0.5, 0.0), i, 3.0, -1.0), i, 3.0, -2.0);
  }
  const int edge[6][2] = {{1, 2}, {2, 1}, {0, 2}, {2, 0}, {0, 1}, {1, 0}};   // (near vertex, far vertex)
  for (int k = 0; k < 6; ++k) phi[3 + k] = mul_lin(mul_lin(mul_lin(one, edge[k][0], 4.5, 0.0), edge[k][1], 1.0, 0.0), edge[k][0], 3.0, -1.0);
  phi[9] = mul_lin(mul_lin(mul_lin(one, 0, 27.0, 0.0), 1, 1.0, 0.0), 2, 1.0, 0.0);
  auto fact = [](int n) { double f = 1.0; for (int k = 2; k <= n; ++k) f *= k; return f; };
  for (int a = 0; a < 10; ++a)
    for (int b = 0; b < 10; ++b) {
      double s = 0.0;
      for (int x = 0; x < 4; ++x) for (int y = 0; y < 4; ++y) for (int z = 0; z < 4; ++z) {
        const double va = phi[a].c[x][y][z];
        if (va == 0.0) continue;
        for (int p = 0; p < 4; ++p) for (int q = 0; q < 4; ++q) for (int r = 0; r < 4; ++r) {
          const double vb = phi[b].c[p][q][r];
          if (vb == 0.0) continue;
          const int ex = x + p, ey = y + q, ez = z + r;
          s += va * vb * 2.0 * fact(ex) * fact(ey) * fact(ez) / fact(ex + ey + ez + 2);
        }
      }
      M[a][b] = s;
    }
}

__global__ void __launch_bounds__(kBlock)
k_l2_error_p3(const int64_t nc, const int32_t* __restrict__ cells, const double2* __restrict__ xy,
              const int32_t* __restrict__ cell_user, const double* __restrict__ uh, const double* __restrict__ table,
              const int local_rows, double* __restrict__ partials, const Fin fin, double* __restrict__ out) {
  __shared__ double red[9];
  __shared__ double sums[1];
  double acc = 0.0;
  for (int64_t c = blockIdx.x * (int64_t)kBlock + threadIdx.x; c < nc; c += (int64_t)gridDim.x * kBlock) {
    const int32_t cu = cell_user[c];
    if (cu < 0) continue;   // counted by the rank that owns the cell's smallest vertex
    const int v0 = cells[3 * c], v1 = cells[3 * c + 1], v2 = cells[3 * c + 2];
    const double2 p0 = xy[v0], p1 = xy[v1], p2 = xy[v2];
    const double area = 0.5 * fabs((p1.x - p0.x) * (p2.y - p0.y) - (p1.y - p0.y) * (p2.x - p0.x));
    const double u0 = uh[v0], u1 = uh[v1], u2 = uh[v2];
    const double t3 = 1.0 / 3.0;
    // the P1 function at the ten P3 nodes (exact embedding), minus the exact solution there
    double e[10] = {u0, u1, u2,
                    t3 * (2.0 * u1 + u2), t3 * (u1 + 2.0 * u2),
                    t3 * (2.0 * u0 + u2), t3 * (u0 + 2.0 * u2),
                    t3 * (2.0 * u0 + u1), t3 * (u0 + 2.0 * u1),
                    t3 * (u0 + u1 + u2)};
    const double* row = table + 10 * (local_rows ? c : (int64_t)cu);
#pragma unroll
    for (int a = 0; a < 10; ++a) e[a] -= row[a];
    double q = 0.0;
#pragma unroll
    for (int a = 0; a < 10; ++a) {
      double s = 0.0;
#pragma unroll
      for (int b = 0; b < 10; ++b) s += kM3[a][b] * e[b];
      q += e[a] * s;
    }
    acc += q * area;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
  Slots<1> sl;
  sl.p[0] = partials;
  if (fin_reduce<1>(fin, sl, gridDim.x, red, sums) && threadIdx.x == 0) out[0] = sums[0];
}

double launch_l2_error_p3(cfem_ctx* c, const double* uh, const double* table, bool local_rows) {
  static bool filled = false;
  static double M[10][10];
  if (!filled) { fill_m3(M); filled = true; }
  CUDA_OK(cudaMemcpyToSymbol(kM3, M, sizeof(M)));   // per call: constant memory is per device, the table is tiny
  if (!fin_available(c)) CFEM_THROW(-1, "l2_error_p3 needs the in-kernel reductions (one GPU or the peer-memory path)");
  const int64_t nc = c->dm.nc;
  int g = (int)std::min<int64_t>((nc + kBlock - 1) / kBlock, (int64_t)c->sm_count * 8);
  if (g < 1) g = 1;
  k_l2_error_p3<<<g, kBlock, 0, c->stream>>>(nc, c->dm.cells, c->dm.xy, c->dm.cell_user, uh, table, local_rows ? 1 : 0,
                                             c->partials + 7 * kMaxPartials, make_fin(c), c->scalars + 25);
  CUDA_OK(cudaGetLastError());
  c->launches.total++;
  CUDA_OK(cudaMemcpyAsync(c->h_pinned + 16, c->scalars + 25, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return c->h_pinned[16];
}

}  // namespace cfem
