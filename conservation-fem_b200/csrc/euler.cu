// 2-D compressible Euler, 4-component P1 system with residual viscosity.
//
// The reference has no working RV Euler solver (Code/Compressible_euler/euler_RV.py is a
// skeleton; only gamma = 1.4 (:33) and the conserved state (rho, m1, m2, E) (:66-72) are
// taken from it).  The scheme is the one stated in oracle/euler.py: group-FEM fluxes,
// BDF2 residual projection, one scalar nodal viscosity, Crank-Nicolson + Newton.
//
// B200 design: with nodally interpolated fluxes the 4x4-block Jacobian never has to be
// stored.  J V = S V + dt/2 (Cx (Ax(U) V) + Cy (Ay(U) V)), with three SCALAR CSR matrices
// sharing the P1 pattern (S = M + dt/2 K_eps changes per step, Cx/Cy = int phi_a d phi_b are
// mesh constants) and nodal 4x4 products done pointwise: 28 B per stored entry instead of
// 132 B for block CSR.  State vectors are AoS (node-major, 4 doubles = one 32-byte gather).
#include <functional>

#include "device_utils.cuh"
#include "launch.h"

namespace cfem {

#define LAUNCHED(c) do { CUDA_OK(cudaGetLastError()); (c)->launches.total++; } while (0)

constexpr double kGamma = 1.4;

static inline int vgrid(const cfem_ctx* c, int64_t n) {
  int64_t b = (n + kBlock - 1) / kBlock;
  const int64_t cap = (int64_t)c->sm_count * 8;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// ---------------------------------------------------------------- nodal (pointwise) kernels
__device__ __forceinline__ void euler_prims(const double4 U, double& u, double& v, double& p) {
  u = U.y / U.x;
  v = U.z / U.x;
  p = (kGamma - 1.0) * (U.w - 0.5 * (U.y * U.y + U.z * U.z) / U.x);
}

__global__ void k_euler_flux(int64_t n, const double4* __restrict__ U, double4* __restrict__ Fx, double4* __restrict__ Fy,
                             double* __restrict__ wave) {
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const double4 q = U[i];
    double u, v, p;
    euler_prims(q, u, v, p);
    if (Fx) {
      Fx[i] = make_double4(q.y, q.y * u + p, q.z * u, (q.w + p) * u);
      Fy[i] = make_double4(q.z, q.y * v, q.z * v + p, (q.w + p) * v);
    }
    if (wave) wave[i] = sqrt(u * u + v * v) + sqrt(kGamma * p / q.x);
  }
}

// Wx = Ax(U) V, Wy = Ay(U) V
__global__ void k_euler_jacvec(int64_t n, const double4* __restrict__ U, const double4* __restrict__ V,
                               double4* __restrict__ Wx, double4* __restrict__ Wy) {
  const double g1 = kGamma - 1.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const double4 q = U[i], d = V[i];
    double u, v, p;
    euler_prims(q, u, v, p);
    const double q2 = u * u + v * v, H = (q.w + p) / q.x;
    Wx[i] = make_double4(
        d.y,
        (0.5 * g1 * q2 - u * u) * d.x + (3.0 - kGamma) * u * d.y - g1 * v * d.z + g1 * d.w,
        -u * v * d.x + v * d.y + u * d.z,
        u * (0.5 * g1 * q2 - H) * d.x + (H - g1 * u * u) * d.y - g1 * u * v * d.z + kGamma * u * d.w);
    Wy[i] = make_double4(
        d.z,
        -u * v * d.x + v * d.y + u * d.z,
        (0.5 * g1 * q2 - v * v) * d.x - g1 * u * d.y + (3.0 - kGamma) * v * d.z + g1 * d.w,
        v * (0.5 * g1 * q2 - H) * d.x - g1 * u * v * d.y + (H - g1 * v * v) * d.z + kGamma * v * d.w);
  }
}

// D = cn Un + co Uold + coo Uoo   (BDF time derivative, nodal)
__global__ void k_lincomb3(int64_t n, double cn, const double* __restrict__ a, double co, const double* __restrict__ b,
                           double coo, const double* __restrict__ c3, double* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    out[i] = cn * a[i] + co * b[i] + coo * c3[i];
}

// ---------------------------------------------------------------- the 4-component CSR apply
struct Apply4 {
  int64_t no;
  const int32_t *rowptr, *colidx;
  const double *A1, *A2, *Cx, *Cy;       // scalar CSR value arrays (nullable)
  const double4 *X1, *X2, *X3, *X4;      // 4-component inputs: A1 X1, A2 X2, Cx X3, Cy X4
  double a1, a2, a3;                     // y = a1 A1 X1 + a2 A2 X2 + a3 (Cx X3 + Cy X4) + add
  const double4* add;                    // nullable
  const uint8_t* is_bc;                  // nullable
  int bc_mode;                           // 0: none, 1: y = X1 (identity row), 2: y = 0, 3: y = X1 - G
  const double4* G;
  double4* y;
};

template <int NDOT>
__global__ void __launch_bounds__(kBlock)
k_apply4(const Apply4 a, const double4* __restrict__ d0, const double4* __restrict__ d1, double* __restrict__ part0,
         double* __restrict__ part1, const int32_t* __restrict__ status) {
  if (status && status[0]) return;
  __shared__ double red[9];
  double acc0 = 0.0, acc1 = 0.0;
  for (int64_t row = blockIdx.x * (int64_t)kBlock + threadIdx.x; row < a.no; row += (int64_t)gridDim.x * kBlock) {
    double4 s = make_double4(0.0, 0.0, 0.0, 0.0);
    const bool bc = a.is_bc && a.is_bc[row];
    if (bc && a.bc_mode != 0) {
      if (a.bc_mode == 1) s = a.X1[row];
      else if (a.bc_mode == 3) { const double4 x = a.X1[row], g = a.G[row]; s = make_double4(x.x - g.x, x.y - g.y, x.z - g.z, x.w - g.w); }
    } else {
      const int p1 = a.rowptr[row + 1];
      for (int p = a.rowptr[row]; p < p1; ++p) {
        const int j = a.colidx[p];
        if (a.A1) { const double w = a.a1 * a.A1[p]; const double4 x = a.X1[j]; s.x += w * x.x; s.y += w * x.y; s.z += w * x.z; s.w += w * x.w; }
        if (a.A2) { const double w = a.a2 * a.A2[p]; const double4 x = a.X2[j]; s.x += w * x.x; s.y += w * x.y; s.z += w * x.z; s.w += w * x.w; }
        if (a.Cx) {
          const double wx = a.a3 * a.Cx[p], wy = a.a3 * a.Cy[p];
          const double4 x = a.X3[j], y = a.X4[j];
          s.x += wx * x.x + wy * y.x; s.y += wx * x.y + wy * y.y; s.z += wx * x.z + wy * y.z; s.w += wx * x.w + wy * y.w;
        }
      }
      if (a.add) { const double4 c0 = a.add[row]; s.x += c0.x; s.y += c0.y; s.z += c0.z; s.w += c0.w; }
    }
    a.y[row] = s;
    if (NDOT >= 1) {
      const double4 d = (d0 == a.y) ? s : d0[row];
      acc0 += s.x * d.x + s.y * d.y + s.z * d.z + s.w * d.w;
    }
    if (NDOT >= 2) {
      const double4 d = (d1 == a.y) ? s : d1[row];
      acc1 += s.x * d.x + s.y * d.y + s.z * d.z + s.w * d.w;
    }
  }
  if (NDOT >= 1) { acc0 = block_sum(acc0, red); if (threadIdx.x == 0) part0[blockIdx.x] = acc0; }
  if (NDOT >= 2) { acc1 = block_sum(acc1, red); if (threadIdx.x == 0) part1[blockIdx.x] = acc1; }
}

static int apply4(cfem_ctx* c, const Apply4& a, int ndot, const double* d0, const double* d1, double* p0, double* p1,
                  bool gated) {
  ProfScope ps(c, PROF_SPMV);
  const int g = vgrid(c, a.no);
  const int32_t* st = gated ? c->status : nullptr;
  if (ndot == 0) k_apply4<0><<<g, kBlock, 0, c->stream>>>(a, nullptr, nullptr, nullptr, nullptr, st);
  else if (ndot == 1) k_apply4<1><<<g, kBlock, 0, c->stream>>>(a, (const double4*)d0, nullptr, p0, nullptr, st);
  else k_apply4<2><<<g, kBlock, 0, c->stream>>>(a, (const double4*)d0, (const double4*)d1, p0, p1, st);
  LAUNCHED(c);
  c->launches.spmv++;
  if (c->world > 1 && ndot >= 1) {
    double* sl[2] = {p0, p1};
    const int op[2] = {0, 0};
    return allreduce_partials(c, ndot, sl, op, g);
  }
  return g;
}

// ---------------------------------------------------------------- Chebyshev mass solve, 4 components
template <bool FIRST>
__global__ void __launch_bounds__(kBlock)
k_cheb4(int64_t no, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const double* __restrict__ vals,
        const double* __restrict__ dinv, const double4* __restrict__ b, const double4* __restrict__ xk,
        double4* __restrict__ xn, double4* __restrict__ d, double c1, double c2, double* __restrict__ part_rr,
        double* __restrict__ part_bb) {
  __shared__ double red[9];
  double rr = 0.0, bb = 0.0;
  for (int64_t row = blockIdx.x * (int64_t)kBlock + threadIdx.x; row < no; row += (int64_t)gridDim.x * kBlock) {
    double4 s = make_double4(0.0, 0.0, 0.0, 0.0);
    const int p1 = rowptr[row + 1];
    for (int p = rowptr[row]; p < p1; ++p) {
      const double w = vals[p];
      const double4 x = xk[colidx[p]];
      s.x += w * x.x; s.y += w * x.y; s.z += w * x.z; s.w += w * x.w;
    }
    const double4 bi = b[row], x0 = xk[row];
    const double di = dinv[row];
    const double4 r = make_double4(bi.x - s.x, bi.y - s.y, bi.z - s.z, bi.w - s.w);
    double4 dk;
    if (FIRST) dk = make_double4(c2 * di * r.x, c2 * di * r.y, c2 * di * r.z, c2 * di * r.w);
    else {
      const double4 dp = d[row];
      dk = make_double4(c1 * dp.x + c2 * di * r.x, c1 * dp.y + c2 * di * r.y, c1 * dp.z + c2 * di * r.z, c1 * dp.w + c2 * di * r.w);
    }
    d[row] = dk;
    xn[row] = make_double4(x0.x + dk.x, x0.y + dk.y, x0.z + dk.z, x0.w + dk.w);
    rr += r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w;
    if (FIRST) bb += bi.x * bi.x + bi.y * bi.y + bi.z * bi.z + bi.w * bi.w;
  }
  rr = block_sum(rr, red);
  if (threadIdx.x == 0) part_rr[blockIdx.x] = rr;
  if (FIRST) { bb = block_sum(bb, red); if (threadIdx.x == 0) part_bb[blockIdx.x] = bb; }
}

__global__ void __launch_bounds__(kBlock)
k_relres4(const double* __restrict__ rrp, int nrr, const double* __restrict__ bbp, int nbb, double* __restrict__ out) {
  __shared__ double red[9];
  const double rr = reduce_partials(rrp, nrr, red);
  const double bb = reduce_partials(bbp, nbb, red);
  if (threadIdx.x == 0) out[0] = bb > 0.0 ? sqrt(rr / bb) : sqrt(rr);
}

static SolveResult chebyshev4(cfem_ctx* c, const Matrix& M, const double* b, double* x, double* tmp, double* d,
                              double rtol, int max_it, int* predict) {
  const int64_t no = c->dm.no, nl = c->dm.nn;
  double *xa = x, *xb = tmp;
  double* prr = c->partials + 3 * kMaxPartials;  // P_RR / P_BB slots of linalg.cu
  double* pbb = c->partials + 4 * kMaxPartials;
  const int g = vgrid(c, no);
  const double theta = 1.25, delta = 0.75, sigma1 = theta / delta;
  double rho = 1.0 / sigma1;
  SolveResult res{0, 0.0, false};
  int it = 0, np_bb = 0;
  int target = predict && *predict > 2 ? *predict : 28;
  if (target > max_it) target = max_it;
  while (true) {
    for (; it < target; ++it) {
      halo_exchange(c, xa, 4);
      ProfScope ps(c, PROF_CHEB);
      if (it == 0) {
        k_cheb4<true><<<g, kBlock, 0, c->stream>>>(no, c->dm.rowptr, c->dm.colidx, M.vals, M.dinv, (const double4*)b,
                                                  (const double4*)xa, (double4*)xb, (double4*)d, 0.0, 1.0 / theta, prr, pbb);
      } else {
        const double rn = 1.0 / (2.0 * sigma1 - rho);
        k_cheb4<false><<<g, kBlock, 0, c->stream>>>(no, c->dm.rowptr, c->dm.colidx, M.vals, M.dinv, (const double4*)b,
                                                   (const double4*)xa, (double4*)xb, (double4*)d, rn * rho, 2.0 * rn / delta, prr, nullptr);
        rho = rn;
      }
      LAUNCHED(c);
      c->launches.spmv++;
      if (it == 0) np_bb = allreduce_sum1(c, pbb, g);
      std::swap(xa, xb);
    }
    const int np_rr = allreduce_sum1(c, prr, g);
    k_relres4<<<1, kBlock, 0, c->stream>>>(prr, np_rr, pbb, np_bb, c->scalars + 7); LAUNCHED(c);
    CUDA_OK(cudaMemcpyAsync(c->h_pinned, c->scalars + 7, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    res.iters = it;
    res.relres = c->h_pinned[0];
    if (!(res.relres == res.relres)) break;
    if (res.relres <= rtol) { res.converged = true; break; }
    if (it >= max_it) break;
    int more = (int)ceil(log(res.relres / rtol) / log(3.0));
    target = it + (more < 2 ? 2 : more);
    if (target > max_it) target = max_it;
  }
  if (xa != x) CUDA_OK(cudaMemcpyAsync(x, xa, 4 * nl * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  halo_exchange(c, x, 4);
  if (predict) {
    int spare = (res.converged && res.relres > 0.0) ? (int)floor(log(rtol / res.relres) / log(3.0)) - 1 : 0;
    if (spare < 0) spare = 0;
    *predict = res.iters - spare > 2 ? res.iters - spare : 2;
  }
  return res;
}

// ---------------------------------------------------------------- viscosity
// partial sum/min/max of the 4 components of U over the owned nodes -> 12 slots starting at `part`
__global__ void __launch_bounds__(kBlock)
k_stats4(int64_t no, const double4* __restrict__ U, double* __restrict__ part) {
  __shared__ double red[9];
  double s[4] = {0, 0, 0, 0}, mn[4] = {INFINITY, INFINITY, INFINITY, INFINITY}, mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < no; i += (int64_t)gridDim.x * kBlock) {
    const double4 q = U[i];
    const double v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) { s[k] += v[k]; mn[k] = fmin(mn[k], v[k]); mx[k] = fmax(mx[k], v[k]); }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double a = block_sum(s[k], red), b = block_min(mn[k], red), m = block_max(mx[k], red);
    if (threadIdx.x == 0) {
      part[(3 * k + 0) * kMaxPartials + blockIdx.x] = a;
      part[(3 * k + 1) * kMaxPartials + blockIdx.x] = b;
      part[(3 * k + 2) * kMaxPartials + blockIdx.x] = m;
    }
  }
}

__global__ void __launch_bounds__(kBlock)
k_eps4(int64_t no, int64_t n_global, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
       const double4* __restrict__ Un, const double4* __restrict__ R, const double* __restrict__ wave,
       const double* __restrict__ h, const double* __restrict__ part, int npart, double Cvel, double Crv,
       double* __restrict__ eps) {
  __shared__ double red[9];
  __shared__ double Ak[4];
  for (int k = 0; k < 4; ++k) {
    double s = 0.0, mn = INFINITY, mx = -INFINITY;
    for (int i = threadIdx.x; i < npart; i += kBlock) {
      s += part[(3 * k + 0) * kMaxPartials + i];
      mn = fmin(mn, part[(3 * k + 1) * kMaxPartials + i]);
      mx = fmax(mx, part[(3 * k + 2) * kMaxPartials + i]);
    }
    s = block_sum(s, red); mn = block_min(mn, red); mx = block_max(mx, red);
    const double mean = s / (double)n_global;
    if (threadIdx.x == 0) Ak[k] = fmax(fabs(mx - mean), fabs(mn - mean));
  }
  __syncthreads();
  for (int64_t row = blockIdx.x * (int64_t)kBlock + threadIdx.x; row < no; row += (int64_t)gridDim.x * kBlock) {
    double umax[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, umin[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
    double rmax[4] = {0, 0, 0, 0}, bmax = 0.0;
    const int p1 = rowptr[row + 1];
    for (int p = rowptr[row]; p < p1; ++p) {
      const int j = colidx[p];
      const double4 u = Un[j], r = R[j];
      const double uv[4] = {u.x, u.y, u.z, u.w}, rv[4] = {fabs(r.x), fabs(r.y), fabs(r.z), fabs(r.w)};
#pragma unroll
      for (int k = 0; k < 4; ++k) { umax[k] = fmax(umax[k], uv[k]); umin[k] = fmin(umin[k], uv[k]); rmax[k] = fmax(rmax[k], rv[k]); }
      bmax = fmax(bmax, wave[j]);
    }
    double Rn = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double nk = fabs((umax[k] - umin[k]) - Ak[k]);
      const double Rk = rmax[k] / nk;
      if (Rk > Rn) Rn = Rk;  // NaN never wins
    }
    const double hi = h[row];
    const double first = __dmul_rn(__dmul_rn(Cvel, hi), bmax);
    const double second = __dmul_rn(__dmul_rn(Crv, __dmul_rn(hi, hi)), fabs(Rn));
    eps[row] = second < first ? second : first;
  }
}

// ---------------------------------------------------------------- data + stepping
struct EulerData {
  double *Uh, *Un, *Uold, *Uoo, *R, *G, *c0, *Fx, *Fy, *Wx, *Wy, *b, *dx, *D, *dinv4, *wave, *tmp, *dch;
  double* work[8];
  Matrix Cx, Cy;
  bool ready = false;
  int cheb_predict = 28, krylov_predict = 8;
};

static EulerData* edata(cfem_ctx* c) {
  if (c->euler) return (EulerData*)c->euler;
  EulerData* e = new EulerData();
  const int64_t nl = c->dm.nn;
  auto alloc = [&](int64_t count) {
    void* p = nullptr;
    CUDA_OK(cudaMalloc(&p, (size_t)count * sizeof(double)));
    CUDA_OK(cudaMemsetAsync(p, 0, (size_t)count * sizeof(double), c->stream));
    c->allocs.push_back(p);
    c->bytes += count * (int64_t)sizeof(double);
    return (double*)p;
  };
  double** v4[] = {&e->Uh, &e->Un, &e->Uold, &e->Uoo, &e->R, &e->G, &e->c0, &e->Fx, &e->Fy, &e->Wx, &e->Wy,
                   &e->b, &e->dx, &e->D, &e->dinv4, &e->tmp, &e->dch};
  for (double** p : v4) *p = alloc(4 * nl);
  for (int k = 0; k < 8; ++k) e->work[k] = alloc(4 * nl);
  e->wave = alloc(nl);
  e->Cx.vals = alloc(c->dm.nnz); e->Cx.dinv = nullptr;
  e->Cy.vals = alloc(c->dm.nnz); e->Cy.dinv = nullptr;
  launch_grad_matrix(c, 0, e->Cx);
  launch_grad_matrix(c, 1, e->Cy);
  e->ready = true;
  c->euler = e;
  return e;
}

__global__ void k_dinv4(int64_t no, const double* __restrict__ dinv, const uint8_t* __restrict__ is_bc, double4* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < no; i += (int64_t)gridDim.x * kBlock) {
    const double d = is_bc[i] ? 1.0 : dinv[i];
    out[i] = make_double4(d, d, d, d);
  }
}
void euler_state_ptrs(cfem_ctx* c, double** Uh, double** Un, double** Uold, double** Uoo, double** G, double** R) {
  EulerData* e = edata(c);
  if (Uh) *Uh = e->Uh;
  if (Un) *Un = e->Un;
  if (Uold) *Uold = e->Uold;
  if (Uoo) *Uoo = e->Uoo;
  if (G) *G = e->G;
  if (R) *R = e->R;
}
void euler_free(cfem_ctx* c) {
  delete (EulerData*)c->euler;  // device buffers belong to ctx->allocs
  c->euler = nullptr;
}
void euler_reset_predictions(cfem_ctx* c) {
  EulerData* e = edata(c);
  e->cheb_predict = 28;
  e->krylov_predict = 8;
}

void euler_steps(cfem_ctx* c, const cfem_step_params* p, int n_steps, cfem_step_stats* st) {
  EulerData* e = edata(c);
  const int64_t no = c->dm.no, nl = c->dm.nn;
  const DevMesh& m = c->dm;
  const double dt = p->dt, hdt = 0.5 * dt;
  Matrix& S = c->mat[CFEM_MAT_SYSTEM];
  Matrix& K = c->mat[CFEM_MAT_STIFFNESS];
  Matrix& M = c->mat[CFEM_MAT_MASS];
  Matrix& Mbc = c->mat[CFEM_MAT_MASS_BC];
  double* normpart = c->partials + 7 * kMaxPartials;
  const int g4 = vgrid(c, 4 * nl), gl = vgrid(c, nl);
  for (int s = 0; s < n_steps; ++s) {
    c->t += dt;
    // ---- residual projection: M_bc R = M D_t U + C.F(U_n), R = 0 on the boundary
    { ProfScope ps(c, PROF_MISC);
      k_euler_flux<<<gl, kBlock, 0, c->stream>>>(nl, (const double4*)e->Un, (double4*)e->Fx, (double4*)e->Fy, nullptr); LAUNCHED(c);
      k_lincomb3<<<g4, kBlock, 0, c->stream>>>(4 * nl, 3.0 / (2.0 * dt), e->Un, -4.0 / (2.0 * dt), e->Uold, 1.0 / (2.0 * dt), e->Uoo, e->D); LAUNCHED(c); }
    Apply4 rhs{no, m.rowptr, m.colidx, M.vals, nullptr, e->Cx.vals, e->Cy.vals, (const double4*)e->D, nullptr,
               (const double4*)e->Fx, (const double4*)e->Fy, 1.0, 0.0, 1.0, nullptr, m.is_bc, 2, nullptr, (double4*)e->b};
    apply4(c, rhs, 0, nullptr, nullptr, nullptr, nullptr, false);
    SolveResult rm = chebyshev4(c, Mbc, e->b, e->R, e->tmp, e->dch, p->lin_rtol, p->lin_max_it, &e->cheb_predict);
    if (!rm.converged) CFEM_THROW(-3, "step_euler: residual mass solve did not converge");
    st->mass_iterations += rm.iters;
    // ---- viscosity
    { ProfScope ps(c, PROF_RV);
      k_euler_flux<<<gl, kBlock, 0, c->stream>>>(nl, (const double4*)e->Uh, nullptr, nullptr, e->wave); LAUNCHED(c);
      const int gs = vgrid(c, no);
      k_stats4<<<gs, kBlock, 0, c->stream>>>(no, (const double4*)e->Uh, c->partials12); LAUNCHED(c);
      int np = gs;
      if (c->world > 1) {
        for (int k = 0; k < 4; ++k) {
          double* sl[3] = {c->partials12 + (3 * k) * kMaxPartials, c->partials12 + (3 * k + 1) * kMaxPartials, c->partials12 + (3 * k + 2) * kMaxPartials};
          const int op[3] = {0, 1, 2};
          np = allreduce_partials(c, 3, sl, op, gs);
        }
      }
      k_eps4<<<gs, kBlock, 0, c->stream>>>(no, m.nn_global, m.rowptr, m.colidx, (const double4*)e->Un, (const double4*)e->R,
                                          e->wave, c->h, c->partials12, np, p->Cvel, p->Crv, c->eps); LAUNCHED(c);
      halo_exchange(c, c->eps); }
    // ---- matrices of the step: K_eps (for the explicit part) and S = M + dt/2 K_eps
    launch_stiffness(c, K, c->eps);
    launch_mass_stiff(c, c->eps, hdt, S);
    { ProfScope ps(c, PROF_MISC); k_dinv4<<<vgrid(c, no), kBlock, 0, c->stream>>>(no, S.dinv, m.is_bc, (double4*)e->dinv4); LAUNCHED(c); }
    // c0 = -M U_n + dt/2 K U_n + dt/2 C.F(U_n)
    Apply4 c0a{no, m.rowptr, m.colidx, M.vals, K.vals, e->Cx.vals, e->Cy.vals, (const double4*)e->Un, (const double4*)e->Un,
               (const double4*)e->Fx, (const double4*)e->Fy, -1.0, hdt, hdt, nullptr, nullptr, 0, nullptr, (double4*)e->c0};
    apply4(c, c0a, 0, nullptr, nullptr, nullptr, nullptr, false);
    // ---- Newton on G(U) = S U + dt/2 C.F(U) + c0
    auto residual = [&]() {
      { ProfScope ps(c, PROF_MISC); k_euler_flux<<<gl, kBlock, 0, c->stream>>>(nl, (const double4*)e->Uh, (double4*)e->Fx, (double4*)e->Fy, nullptr); LAUNCHED(c); }
      Apply4 ga{no, m.rowptr, m.colidx, S.vals, nullptr, e->Cx.vals, e->Cy.vals, (const double4*)e->Uh, nullptr,
                (const double4*)e->Fx, (const double4*)e->Fy, 1.0, 0.0, hdt, (const double4*)e->c0, m.is_bc, 3,
                (const double4*)e->G, (double4*)e->b};
      int np = apply4(c, ga, 1, e->b, nullptr, normpart, nullptr, false);
      double* tmp = c->h_pinned + 64;
      CUDA_OK(cudaMemcpyAsync(tmp, normpart, np * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CUDA_OK(cudaStreamSynchronize(c->stream));
      double sum = 0.0;
      for (int i = 0; i < np; ++i) sum += tmp[i];
      return sqrt(sum);
    };
    double res = residual();
    const double res0 = res;
    bool converged = res < p->newton_atol;
    int it = 0;
    while (!converged && it < p->newton_max_it) {
      // J V = S V + dt/2 (Cx (Ax(Uh) V) + Cy (Ay(Uh) V)), identity on Dirichlet rows
      LinApply op = [&](const double* x, double* y, int ndot, const double* d0, const double* d1, double* p0, double* p1, bool gated) {
        halo_exchange(c, const_cast<double*>(x), 4);
        { ProfScope ps(c, PROF_KRYLOV_VEC);
          k_euler_jacvec<<<gl, kBlock, 0, c->stream>>>(nl, (const double4*)e->Uh, (const double4*)x, (double4*)e->Wx, (double4*)e->Wy); LAUNCHED(c); }
        Apply4 ja{no, m.rowptr, m.colidx, S.vals, nullptr, e->Cx.vals, e->Cy.vals, (const double4*)x, nullptr,
                  (const double4*)e->Wx, (const double4*)e->Wy, 1.0, 0.0, hdt, nullptr, m.is_bc, 1, nullptr, (double4*)y};
        return apply4(c, ja, ndot, d0, d1, p0, p1, gated);
      };
      CUDA_OK(cudaMemsetAsync(e->dx, 0, 4 * nl * sizeof(double), c->stream));
      SolveResult rk = bicgstab_generic(c, 4 * no, 4, e->dinv4, op, e->work, e->b, e->dx, p->lin_rtol, 0.0, p->lin_max_it,
                                        &e->krylov_predict);
      if (!rk.converged) CFEM_THROW(-3, "step_euler: Krylov solve did not converge (relres " + std::to_string(rk.relres) + ")");
      st->krylov_iterations += rk.iters;
      launch_sub(c, e->Uh, e->dx, 4 * no);
      halo_exchange(c, e->Uh, 4);
      ++it;
      res = residual();
      converged = (res / res0 < p->newton_rtol) || (res < p->newton_atol);
    }
    st->newton_iterations += it;
    st->last_newton_residual = res;
    if (!converged) CFEM_THROW(-3, "Euler Newton solver did not converge in " + std::to_string(it) + " iterations");
    double* t = e->Uoo;
    e->Uoo = e->Uold;
    e->Uold = e->Un;
    e->Un = t;
    CUDA_OK(cudaMemcpyAsync(e->Un, e->Uh, 4 * nl * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    st->steps++;
  }
}

}  // namespace cfem
